"""torch-facing wrappers (autograd Functions) over the C ABI.  torch is used for device memory,
streams and autograd bookkeeping only; every FLOP below runs in the hand-written CUDA library.
"""
import torch

from cap2det_b200 import capi
from cap2det_b200.capi import call, ptr, stream, require_cuda

HEAD_FEATURE_DIMS = 1024
HEAD_IN_HW = 7
HEAD_IN_CH = 576


def _rows(t):
  """[..., n] float32 view whose rows are `ld` floats apart -> (ptr, ld, rows, n)."""
  import ctypes
  if t.dtype != torch.float32 or not t.is_cuda:
    raise ValueError('expected a CUDA float32 tensor')
  n = t.shape[-1]
  rows = t.numel() // max(n, 1)
  if t.is_contiguous():
    return ctypes.c_void_p(t.data_ptr()), n, rows, n
  if t.stride(-1) != 1:
    raise ValueError('last dimension must be dense')
  ld = t.stride(-2)
  expect = ld
  for d in range(t.dim() - 2, -1, -1):
    if t.shape[d] != 1 and t.stride(d) != expect:
      raise ValueError('rows must be uniformly strided')
    expect *= t.shape[d]
  return ctypes.c_void_p(t.data_ptr()), ld, rows, n


def _f32(t):
  if t.dtype != torch.float32:
    raise ValueError('expected float32 tensor, got %s' % t.dtype)
  return t


# ---------------------------------------------------------------------------------------------
# K1: crop_and_resize + max_pool  (models/utils.py:147-160)
# ---------------------------------------------------------------------------------------------
class PoolFold(object):
  """Links one roi_crop_maxpool call with the head_mixed5 call that consumes its output (bf16 path, training): the
  head's backward then leaves the backward of its first max-pool (Mixed_5a/Branch_2) out of dx0 and hands the
  pool's arg-max codes and output gradient to the ROI backward, which applies that term while it scatters
  (c2d_head_mixed5_bwd_fold / c2d_roi_crop_maxpool_bwd_codes_fold).  Pass the SAME object to both calls."""

  def __init__(self):
    self.ws = None            # keeps the head workspace alive until the ROI backward has run
    self.pool_codes = None
    self.pool_grad = None
    self.pool_grad_ld = 0

  @property
  def ready(self):
    return self.ws is not None


class _RoiCropMaxPool(torch.autograd.Function):
  """When the feature map needs a gradient the forward also writes the max-pool arg-max codes (1 byte per
  bin and channel quad) and the backward scatters from them, instead of re-sampling the feature map."""

  @staticmethod
  def forward(ctx, fmap, proposals, crop_size, pool_k, pool_s, out_dtype, fold):
    require_cuda(fmap, proposals)
    ctx.fold = fold
    _f32(fmap); _f32(proposals)
    B, Hf, Wf, Cf = fmap.shape
    P = proposals.shape[1]
    hp = crop_size // pool_s
    out = torch.empty((B * P, hp, hp, Cf), dtype=out_dtype, device=fmap.device)
    codes = None
    if ctx.needs_input_grad[0] and B * P > 0:
      codes = torch.empty((capi.load().c2d_roi_argmax_code_bytes(B * P, Cf, crop_size),), dtype=torch.uint8,
                          device=fmap.device)
    call('c2d_roi_crop_maxpool_fwd_codes', ptr(fmap), B, Hf, Wf, Cf, ptr(proposals), P, crop_size, pool_k, pool_s,
         ptr(out), capi.dtype_code(out_dtype), ptr(codes), stream())
    ctx.save_for_backward(proposals, codes)
    ctx.cfg = (crop_size, pool_k, pool_s)
    ctx.fmap_shape = (B, Hf, Wf, Cf)
    return out

  @staticmethod
  def backward(ctx, dout):
    proposals, codes = ctx.saved_tensors
    crop_size, pool_k, pool_s = ctx.cfg
    B, Hf, Wf, Cf = ctx.fmap_shape
    P = proposals.shape[1]
    dout = dout.contiguous()
    dfmap = torch.empty((B, Hf, Wf, Cf), dtype=torch.float32, device=dout.device)
    fold = ctx.fold
    folded = fold is not None and fold.ready
    ws_bytes = capi.load().c2d_roi_bwd_tiles_workspace_bytes(B, Hf, Wf, Cf, P, crop_size, 1 if folded else 0)
    if ws_bytes:
      # tile-owner backward: gradients are summed across proposals on chip (csrc/c2d_roi.cu)
      ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dout.device)
      call('c2d_roi_crop_maxpool_bwd_tiles', B, Hf, Wf, Cf, ptr(proposals), P, crop_size, pool_k, pool_s, ptr(codes),
           ptr(dout), capi.dtype_code(dout.dtype), fold.pool_codes if folded else None,
           fold.pool_grad if folded else None, fold.pool_grad_ld if folded else 0, ptr(ws), ws_bytes, ptr(dfmap), stream())
    elif folded:
      call('c2d_roi_crop_maxpool_bwd_codes_fold', B, Hf, Wf, Cf, ptr(proposals), P, crop_size, pool_k, pool_s, ptr(codes),
           ptr(dout), fold.pool_codes, fold.pool_grad, fold.pool_grad_ld, ptr(dfmap), stream())
    else:
      call('c2d_roi_crop_maxpool_bwd_codes', B, Hf, Wf, Cf, ptr(proposals), P, crop_size, pool_k, pool_s, ptr(codes),
           ptr(dout), capi.dtype_code(dout.dtype), ptr(dfmap), stream())
    if folded:
      fold.ws = None
    return dfmap, None, None, None, None, None, None


def roi_crop_maxpool(fmap, proposals, crop_size=14, pool_k=2, pool_s=2, out_dtype=torch.float32, fold=None):
  """fmap [B,Hf,Wf,C] fp32, proposals [B,P,4] normalised -> [B*P, crop/2, crop/2, C].  `fold`: see PoolFold."""
  return _RoiCropMaxPool.apply(fmap.contiguous(), proposals.contiguous(), crop_size, pool_k, pool_s, out_dtype, fold)


def roi_crop_maxpool_multi(fmaps, proposals, crop_size=14, pool_k=2, pool_s=2, out_dtype=torch.float32):
  """Inference only: the SAME proposals [1,P,4] cropped from several feature maps [1,Hf_s,Wf_s,C] of different sizes
  (multi-scale evaluation, models/cap2det_model.py:231-272) into ONE tensor [S*P, crop/2, crop/2, C], scale-major -- so
  that the head, the FC layers and MIDN run once over all scales as a batch of S.  Every slice is exactly what
  roi_crop_maxpool returns for that feature map."""
  require_cuda(proposals, *fmaps)
  _f32(proposals)
  proposals = proposals.contiguous()
  if proposals.shape[0] != 1:
    raise ValueError('roi_crop_maxpool_multi takes one image (proposals [1,P,4])')
  P = proposals.shape[1]
  hp = crop_size // pool_s
  Cf = fmaps[0].shape[-1]
  out = torch.empty((len(fmaps) * P, hp, hp, Cf), dtype=out_dtype, device=proposals.device)
  for s, fmap in enumerate(fmaps):
    _f32(fmap)
    fmap = fmap.contiguous()
    if fmap.shape[0] != 1 or fmap.shape[-1] != Cf:
      raise ValueError('feature maps must be [1,Hf,Wf,%d]' % Cf)
    call('c2d_roi_crop_maxpool_fwd_codes', ptr(fmap), 1, fmap.shape[1], fmap.shape[2], Cf, ptr(proposals), P, crop_size,
         pool_k, pool_s, ptr(out[s * P:(s + 1) * P]), capi.dtype_code(out_dtype), None, stream())
  return out


# ---------------------------------------------------------------------------------------------
# K2/K3: Mixed_5a..5c head + spatial mean + dropout  (models/utils.py:165-177)
# ---------------------------------------------------------------------------------------------
def head_conv_specs():
  """[(tf_scope, k, cin, cout, stride, offsets dict)] of the packed head parameter buffer."""
  import ctypes
  lib = capi.load()
  out = []
  for i in range(lib.c2d_head_num_convs()):
    k, cin, cout, stride = (ctypes.c_int() for _ in range(4))
    name = ctypes.c_char_p()
    capi.check(lib.c2d_head_conv_spec(i, ctypes.byref(k), ctypes.byref(cin), ctypes.byref(cout),
                                      ctypes.byref(stride), ctypes.byref(name)))
    offs = [ctypes.c_longlong() for _ in range(5)]
    capi.check(lib.c2d_head_param_offsets(i, *[ctypes.byref(o) for o in offs]))
    out.append((name.value.decode(), k.value, cin.value, cout.value, stride.value,
                dict(zip(('weights', 'gamma', 'beta', 'moving_mean', 'moving_variance'),
                         [o.value for o in offs]))))
  return out


def head_param_floats():
  return int(capi.load().c2d_head_param_floats())


class _HeadMixed5(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x0, params, keep_mask, keep_prob, need_dx0, fold):
    require_cuda(x0, params, keep_mask)
    ctx.fold = fold
    _f32(params)
    n = x0.shape[0]
    if tuple(x0.shape[1:]) != (HEAD_IN_HW, HEAD_IN_HW, HEAD_IN_CH):
      raise ValueError('head expects [N,7,7,576] ROI features, got %s' % (tuple(x0.shape),))
    dt = capi.dtype_code(x0.dtype)
    lib = capi.load()
    nbytes = lib.c2d_head_workspace_bytes(n, dt)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=x0.device)
    feat = torch.empty((n, HEAD_FEATURE_DIMS), dtype=torch.float32, device=x0.device)
    call('c2d_head_mixed5_fwd', ptr(x0), n, dt, ptr(params), ptr(ws), nbytes, ptr(keep_mask), float(keep_prob),
         ptr(feat), stream())
    ctx.save_for_backward(x0, params, keep_mask, ws)
    ctx.keep_prob = float(keep_prob)
    ctx.need_dx0 = need_dx0
    return feat

  @staticmethod
  def backward(ctx, dfeat):
    x0, params, keep_mask, ws = ctx.saved_tensors
    n = x0.shape[0]
    dt = capi.dtype_code(x0.dtype)
    dfeat = dfeat.contiguous()
    dparams = torch.empty_like(params)
    dx0 = torch.empty_like(x0) if (ctx.need_dx0 and ctx.needs_input_grad[0]) else None
    fold = ctx.fold
    if fold is not None and dx0 is not None and x0.dtype == torch.bfloat16:
      import ctypes
      codes, grad, ld = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int()
      call('c2d_head_mixed5_bwd_fold', ptr(x0), n, dt, ptr(params), ptr(ws), ws.numel(), ptr(keep_mask), ctx.keep_prob,
           ptr(dfeat), ptr(dparams), ptr(dx0), ctypes.byref(codes), ctypes.byref(grad), ctypes.byref(ld), stream())
      fold.ws, fold.pool_codes, fold.pool_grad, fold.pool_grad_ld = ws, codes, grad, ld.value
    else:
      call('c2d_head_mixed5_bwd', ptr(x0), n, dt, ptr(params), ptr(ws), ws.numel(), ptr(keep_mask), ctx.keep_prob,
           ptr(dfeat), ptr(dparams), ptr(dx0), stream())
    return dx0, dparams, None, None, None, None


def head_mixed5(x0, params, keep_mask=None, keep_prob=1.0, need_dx0=True, fold=None):
  """x0 [N,7,7,576] (fp32 or bf16), packed params fp32 -> proposal features [N,1024] fp32.  `fold`: see PoolFold
  (only when x0 is exactly the tensor roi_crop_maxpool returned for the same PoolFold object)."""
  return _HeadMixed5.apply(x0.contiguous(), params, keep_mask, keep_prob, need_dx0, fold)


# ---------------------------------------------------------------------------------------------
# Reader-side image / box contract  (readers/cap2det_reader.py:143-199, core/imgproc.py:300-352)
# ---------------------------------------------------------------------------------------------
class _DropoutApply(torch.autograd.Function):
  @staticmethod
  def forward(ctx, x, mask, keep_prob):
    require_cuda(x, mask)
    _f32(x); _f32(mask)
    if mask.shape != x.shape:
      raise ValueError('keep mask %s does not match the tensor %s' % (tuple(mask.shape), tuple(x.shape)))
    out = torch.empty_like(x)
    call('c2d_dropout_apply', ptr(x), ptr(mask), float(keep_prob), ptr(out), x.numel(), stream())
    ctx.save_for_backward(mask)
    ctx.keep_prob = float(keep_prob)
    return out

  @staticmethod
  def backward(ctx, dout):
    (mask,) = ctx.saved_tensors
    dout = dout.contiguous()
    dx = torch.empty_like(dout)
    call('c2d_dropout_apply', ptr(dout), ptr(mask), ctx.keep_prob, ptr(dx), dout.numel(), stream())
    return dx, None, None


def dropout_apply(x, keep_mask, keep_prob):
  """slim.dropout with a given {0,1} keep mask: (x / keep_prob) * mask (tf.nn.dropout of TF 1.x), differentiable in x.
  Used for frcnn_options.dropout_on_feature_map (models/utils.py:138-142)."""
  return _DropoutApply.apply(x.contiguous(), keep_mask.contiguous(), keep_prob)


def dropout_keep_mask(state, seed, shape, keep_prob):
  """slim.dropout's keep mask floor(keep_prob + uniform[0,1)) (models/utils.py:176-177) from the library's Philox kernel.
  state: int64[2] device tensor, zero-initialised once and then owned by the kernel (it counts the masks drawn)."""
  require_cuda(state)
  mask = torch.empty(tuple(shape), dtype=torch.float32, device=state.device)
  call('c2d_dropout_keep_mask', ptr(state), int(seed) & 0xffffffff, mask.numel(), float(keep_prob), ptr(mask), stream())
  return mask


def resize_bilinear(image, new_height, new_width):
  """tf.image.resize_images(image, [new_height, new_width]) (bilinear, align_corners False, TF1 sampling).
  image [B,H,W,C] or [H,W,C], fp32 or uint8 -> fp32."""
  squeeze = image.dim() == 3
  x = (image.unsqueeze(0) if squeeze else image).contiguous()
  require_cuda(x)
  B, H, W, C = x.shape
  out = torch.empty((B, int(new_height), int(new_width), C), dtype=torch.float32, device=x.device)
  call('c2d_resize_bilinear', ptr(x), capi.dtype_code(x.dtype), B, H, W, C, ptr(out), int(new_height), int(new_width),
       stream())
  return out[0] if squeeze else out


def image_flip_left_right(image, flip=None):
  """tf.image.flip_left_right; ``flip`` [B] selects the images to flip (None = all)."""
  squeeze = image.dim() == 3
  x = (image.unsqueeze(0) if squeeze else image).contiguous()
  f = None if flip is None else flip.to(torch.int32).contiguous()
  require_cuda(x, f)
  B, H, W, C = x.shape
  out = torch.empty_like(x)
  call('c2d_image_flip_left_right', ptr(x), capi.dtype_code(x.dtype), B, H, W, C, ptr(f), ptr(out), stream())
  return out[0] if squeeze else out


def box_scale_batch(box, image_hw, pad_h, pad_w):
  """_batch_scale_box_fn: box [B,P,4] * image_hw[b] / (pad_h, pad_w)."""
  box = _f32(box).contiguous()
  hw = image_hw.to(torch.int32).contiguous()
  require_cuda(box, hw)
  B, P, _ = box.shape
  out = torch.empty_like(box)
  call('c2d_box_scale_batch', ptr(box), ptr(hw), B, P, int(pad_h), int(pad_w), ptr(out), stream())
  return out


# ---------------------------------------------------------------------------------------------
# First stage: Inception-v2 up to Mixed_4e on whole images  (models/utils.py:127-136)
# ---------------------------------------------------------------------------------------------
BACKBONE_STEM_SCOPE = 'Conv2d_1a_7x7'


def backbone_conv_specs():
  """[(tf_scope, k, cin, cout, stride, offsets dict)] of the packed first-stage parameter buffer; the
  separable stem comes first with offsets 'depthwise_weights' [7,7,3,8] and 'pointwise_weights' [64,24]."""
  import ctypes
  lib = capi.load()
  names = ('weights', 'gamma', 'beta', 'moving_mean', 'moving_variance')
  offs = [ctypes.c_longlong() for _ in range(5)]
  capi.check(lib.c2d_backbone_param_offsets(-1, *[ctypes.byref(o) for o in offs]))
  stem = dict(zip(names, [o.value for o in offs]))
  stem['depthwise_weights'] = stem.pop('weights')
  stem['pointwise_weights'] = stem['depthwise_weights'] + 7 * 7 * 3 * 8
  out = [(BACKBONE_STEM_SCOPE, 7, 3, 64, 2, stem)]
  for i in range(lib.c2d_backbone_num_convs()):
    k, cin, cout, stride = (ctypes.c_int() for _ in range(4))
    name = ctypes.c_char_p()
    capi.check(lib.c2d_backbone_conv_spec(i, ctypes.byref(k), ctypes.byref(cin), ctypes.byref(cout),
                                          ctypes.byref(stride), ctypes.byref(name)))
    capi.check(lib.c2d_backbone_param_offsets(i, *[ctypes.byref(o) for o in offs]))
    out.append((name.value.decode(), k.value, cin.value, cout.value, stride.value,
                dict(zip(names, [o.value for o in offs]))))
  return out


def backbone_param_floats():
  return int(capi.load().c2d_backbone_param_floats())


def backbone_out_dims(height, width):
  import ctypes
  hf, wf = ctypes.c_int(), ctypes.c_int()
  capi.check(capi.load().c2d_backbone_out_dims(int(height), int(width), ctypes.byref(hf), ctypes.byref(wf)))
  return hf.value, wf.value


class _BackboneInceptionV2(torch.autograd.Function):

  @staticmethod
  def forward(ctx, image, params):
    require_cuda(image, params)
    _f32(image); _f32(params)
    B, H, W, C = image.shape
    if C != 3:
      raise ValueError('backbone expects [B,H,W,3] images, got %s' % (tuple(image.shape),))
    Hf, Wf = backbone_out_dims(H, W)
    nbytes = capi.load().c2d_backbone_workspace_bytes(B, H, W)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=image.device)
    fmap = torch.empty((B, Hf, Wf, BACKBONE_OUT_CH), dtype=torch.float32, device=image.device)
    call('c2d_backbone_fwd', ptr(image), B, H, W, ptr(params), ptr(ws), nbytes, ptr(fmap), stream())
    ctx.save_for_backward(params, ws, fmap)
    ctx.shape = (B, H, W)
    return fmap

  @staticmethod
  def backward(ctx, dfmap):
    params, ws, fmap = ctx.saved_tensors
    B, H, W = ctx.shape
    dparams = torch.empty_like(params)
    dfmap = dfmap.contiguous()              # named: ptr() of a temporary would dangle
    call('c2d_backbone_bwd', ptr(dfmap), ptr(fmap), B, H, W, ptr(params), ptr(ws), ws.numel(),
         ptr(dparams), stream())
    return None, dparams


BACKBONE_OUT_CH = 576


def backbone_mixed4e_input(fmap):
  """Parity-test hook: the Mixed_4e input (bf16 [B,Hf,Wf,576]) of the forward pass that produced ``fmap``."""
  fn = fmap.grad_fn
  if fn is None or not hasattr(fn, 'saved_tensors'):
    raise ValueError('fmap must come from backbone_inception_v2 with a params tensor that requires grad')
  _, ws, _ = fn.saved_tensors
  B, H, W = fn.shape
  x = torch.empty(tuple(fmap.shape), dtype=torch.bfloat16, device=fmap.device)
  call('c2d_backbone_mixed4e_input', ptr(ws), B, H, W, ptr(x), stream())
  return x


def backbone_inception_v2(image, params):
  """image [B,H,W,3] fp32 pixel values in [0,255], packed first-stage params fp32 -> features_to_crop
  [B,ceil(H/16),ceil(W/16),576] fp32.  Backward yields the Mixed_4e variable gradients only."""
  return _BackboneInceptionV2.apply(image.contiguous(), params)


# ---------------------------------------------------------------------------------------------
# K4: concatenated fully connected layers  (models/cap2det_model.py:79-88,190-197)
# ---------------------------------------------------------------------------------------------
def _ld16(n):
  return (n + 15) // 16 * 16


class _FcConcat(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x, w, b, dtype_code):
    require_cuda(x, w, b)
    M, D = x.shape
    N = w.shape[0]
    ld = _ld16(N)
    # the tensor-core path writes all ld columns (padding = zero weights + zero bias); the CUDA-core path only the first N
    y = (torch.empty if dtype_code == capi.C2D_BF16 else torch.zeros)((M, ld), dtype=torch.float32, device=x.device)
    ws, nbytes = None, 0
    if dtype_code == capi.C2D_BF16:
      nbytes = capi.load().c2d_fc_workspace_bytes(M, D, N, dtype_code)
      ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
    call('c2d_fc_fwd', ptr(x), M, D, ptr(w), ptr(b), N, ptr(y), ld, dtype_code, ptr(ws), nbytes, stream())
    ctx.save_for_backward(x, w)
    ctx.dtype_code = dtype_code
    return y

  @staticmethod
  def backward(ctx, dy):
    x, w = ctx.saved_tensors
    M, D = x.shape
    N = w.shape[0]
    ld = _ld16(N)
    dy = dy.contiguous()
    lib = capi.load()
    nbytes = lib.c2d_fc_workspace_bytes(M, D, N, ctx.dtype_code)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
    dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
    dw = torch.empty_like(w)
    db = torch.empty((N,), dtype=torch.float32, device=x.device)
    call('c2d_fc_bwd', ptr(x), M, D, ptr(w), N, ptr(dy), ld, ptr(dx), ptr(dw), ptr(db), ctx.dtype_code, ptr(ws),
         nbytes, stream())
    return dx, dw, db, None


def fc_concat(x, w, b, compute_dtype=torch.float32):
  """x [M,D] . w [N,D]^T + b [N] -> y [M, ld16(N)] fp32 (columns >= N are zero padding).
  compute_dtype bfloat16 runs the product on the tensor cores (bf16 operands, fp32 accumulation)."""
  return _FcConcat.apply(x.contiguous(), w.contiguous(), b.contiguous(), capi.dtype_code(compute_dtype))


# ---------------------------------------------------------------------------------------------
# K5: MIDN  (models/cap2det_model.py:70-109)
# ---------------------------------------------------------------------------------------------
class _Midn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, logits_all, col_r, col_c, num_classes, num_proposals):
    """logits_all [B,P,ld] holds both streams: columns [col_r, col_r+C) and [col_c, col_c+C)."""
    require_cuda(logits_all, num_proposals)
    B, P, ld = logits_all.shape
    C = num_classes
    dev = logits_all.device
    class_logits = torch.empty((B, C), dtype=torch.float32, device=dev)
    scores = torch.empty((B, P, C), dtype=torch.float32, device=dev)
    proba = torch.empty((B, P, C), dtype=torch.float32, device=dev)
    base = logits_all.data_ptr()
    import ctypes
    call('c2d_midn_fwd', ctypes.c_void_p(base + 4 * col_r), ctypes.c_void_p(base + 4 * col_c), ld,
         ptr(num_proposals), B, P, C, ptr(class_logits), ptr(scores), ptr(proba), stream())
    ctx.save_for_backward(logits_all, num_proposals, class_logits, proba)
    ctx.cfg = (col_r, col_c, C)
    ctx.set_materialize_grads(False)
    return class_logits, scores, proba

  @staticmethod
  def backward(ctx, d_cl, d_sc, d_pr):
    import ctypes
    logits_all, num_proposals, class_logits, proba = ctx.saved_tensors
    col_r, col_c, C = ctx.cfg
    B, P, ld = logits_all.shape
    d_all = torch.zeros_like(logits_all)
    d_cl = d_cl.contiguous() if d_cl is not None else None
    d_sc = d_sc.contiguous() if d_sc is not None else None
    d_pr = d_pr.contiguous() if d_pr is not None else None
    base = logits_all.data_ptr()
    dbase = d_all.data_ptr()
    call('c2d_midn_bwd', ctypes.c_void_p(base + 4 * col_c), ld, ptr(num_proposals), B, P, C, ptr(class_logits),
         ptr(proba), ptr(d_cl), ptr(d_sc), ptr(d_pr), ctypes.c_void_p(dbase + 4 * col_r),
         ctypes.c_void_p(dbase + 4 * col_c), ld, stream())
    return d_all, None, None, None, None


def midn(logits_all, col_r, col_c, num_classes, num_proposals):
  return _Midn.apply(logits_all, col_r, col_c, num_classes, num_proposals)


# ---------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------
class _SigmoidCeMean(torch.autograd.Function):

  @staticmethod
  def forward(ctx, labels, logits, weight):
    require_cuda(labels, logits)
    loss = torch.empty((), dtype=torch.float32, device=logits.device)
    call('c2d_sigmoid_ce_mean_fwd', ptr(labels), ptr(logits), logits.numel(), float(weight), ptr(loss), stream())
    ctx.save_for_backward(labels, logits)
    ctx.weight = float(weight)
    return loss

  @staticmethod
  def backward(ctx, dloss):
    labels, logits = ctx.saved_tensors
    dlogits = torch.empty_like(logits)
    dloss = dloss.contiguous()
    call('c2d_sigmoid_ce_mean_bwd', ptr(labels), ptr(logits), logits.numel(), ctx.weight, ptr(dloss), ptr(dlogits),
         stream())
    return None, dlogits, None


def sigmoid_ce_mean(labels, logits, weight=1.0):
  """weight * reduce_mean(sigmoid_cross_entropy_with_logits)  (models/cap2det_model.py:293-297)."""
  return _SigmoidCeMean.apply(labels.contiguous(), logits.contiguous(), weight)


def softmax_rows(x):
  """tf.nn.softmax(x, axis=-1) for a [..., n] tensor (may be a column slice of a wider buffer)."""
  p, ld, rows, n = _rows(x)
  y = torch.empty(tuple(x.shape), dtype=torch.float32, device=x.device)
  call('c2d_softmax_rows', p, ld, rows, n, ptr(y), n, stream())
  return y


def oicr_assign(labels, num_proposals, proposals, scores0_cls, iou_threshold):
  """models/utils.py:37-95.  scores0_cls [B,P,C] (class columns of the previous stage; may be a slice).

  Returns (proposal_ind [B,C] int64, proposal_labels [B,P,1+C], status int32 tensor)."""
  require_cuda(labels, num_proposals, proposals)
  B, P, C = scores0_cls.shape
  s0_ptr, ld0, _, _ = _rows(scores0_cls)
  dev = labels.device
  ind = torch.empty((B, C), dtype=torch.int64, device=dev)
  pl = torch.empty((B, P, C + 1), dtype=torch.float32, device=dev)
  status = torch.empty((1,), dtype=torch.int32, device=dev)
  call('c2d_oicr_assign', ptr(labels), ptr(num_proposals), ptr(proposals), s0_ptr, ld0,
       float(iou_threshold), B, P, C, ptr(ind), ptr(pl), ptr(status), stream())
  return ind, pl, status


class _OicrCe(torch.autograd.Function):

  @staticmethod
  def forward(ctx, logits_all, col, proposal_labels, num_proposals, weight):
    import ctypes
    require_cuda(logits_all, proposal_labels, num_proposals)
    B, P, ld = logits_all.shape
    C = proposal_labels.shape[-1] - 1
    loss = torch.empty((), dtype=torch.float32, device=logits_all.device)
    call('c2d_oicr_ce_fwd', ptr(proposal_labels), ctypes.c_void_p(logits_all.data_ptr() + 4 * col), ld,
         ptr(num_proposals), B, P, C, float(weight), ptr(loss), stream())
    ctx.save_for_backward(logits_all, proposal_labels, num_proposals)
    ctx.cfg = (col, float(weight))
    return loss

  @staticmethod
  def backward(ctx, dloss):
    import ctypes
    logits_all, pl, num_proposals = ctx.saved_tensors
    col, weight = ctx.cfg
    B, P, ld = logits_all.shape
    C = pl.shape[-1] - 1
    d_all = torch.zeros_like(logits_all)
    dloss = dloss.contiguous()
    call('c2d_oicr_ce_bwd', ptr(pl), ctypes.c_void_p(logits_all.data_ptr() + 4 * col), ld, ptr(num_proposals),
         B, P, C, weight, ptr(dloss), ctypes.c_void_p(d_all.data_ptr() + 4 * col), ld, stream())
    return d_all, None, None, None, None


def oicr_cross_entropy(logits_all, col, proposal_labels, num_proposals, weight=1.0):
  """weight * calc_oicr_loss's soft-label CE on columns [col, col+1+C) of logits_all."""
  return _OicrCe.apply(logits_all, col, proposal_labels.contiguous(), num_proposals, weight)


class _LossHead(torch.autograd.Function):
  """models/cap2det_model.py:274-330 as ONE autograd node over the concatenated logits (c2d_loss_head_fwd / _bwd)."""

  @staticmethod
  def forward(ctx, logits_all, labels, num_proposals, proposals, class_logits, proba, scores0, cfg):
    require_cuda(logits_all, labels, num_proposals, proposals, class_logits, proba, scores0)
    C, K, col_r, col_c, col_oicr0, thr, w_midn, w_oicr = cfg
    B, P, ld = logits_all.shape
    dev = logits_all.device
    losses = torch.empty((K + 2,), dtype=torch.float32, device=dev)
    ind = torch.empty((max(K, 1), B, C), dtype=torch.int64, device=dev)
    pl = torch.empty((max(K, 1), B, P, C + 1), dtype=torch.float32, device=dev)
    sm = torch.empty((max(K - 1, 1), B, P, C + 1), dtype=torch.float32, device=dev)
    status = torch.empty((1,), dtype=torch.int32, device=dev)
    call('c2d_loss_head_fwd', ptr(logits_all), ld, ptr(num_proposals), ptr(proposals), ptr(labels), ptr(class_logits),
         ptr(scores0), B, P, C, K, col_oicr0, float(thr), float(w_midn), float(w_oicr), ptr(sm), ptr(ind), ptr(pl),
         ptr(losses), ptr(status), stream())
    ctx.save_for_backward(logits_all, labels, num_proposals, class_logits, proba, pl)
    ctx.cfg = cfg
    ctx.set_materialize_grads(False)
    outs = tuple(losses[i] for i in range(K + 2))
    ctx.mark_non_differentiable(ind, pl, status)
    return outs + (ind, pl, status)

  @staticmethod
  def backward(ctx, *grads):
    logits_all, labels, num_proposals, class_logits, proba, pl = ctx.saved_tensors
    C, K, col_r, col_c, col_oicr0, thr, w_midn, w_oicr = ctx.cfg
    B, P, ld = logits_all.shape
    g = [x.contiguous() if x is not None else None for x in grads[:K + 2]]
    d_oicr = [g[1 + k] if k < K else None for k in range(4)]
    d_all = torch.empty_like(logits_all)
    call('c2d_loss_head_bwd', ptr(logits_all), ld, ptr(num_proposals), ptr(labels), ptr(class_logits), ptr(proba), ptr(pl),
         B, P, C, K, col_r, col_c, col_oicr0, float(w_midn), float(w_oicr), ptr(g[0]), ptr(d_oicr[0]), ptr(d_oicr[1]),
         ptr(d_oicr[2]), ptr(d_oicr[3]), ptr(g[K + 1]), ptr(d_all), stream())
    return d_all, None, None, None, None, None, None, None


def loss_head(logits_all, labels, num_proposals, proposals, class_logits, proba, scores0, num_classes, num_stages, col_r,
              col_c, col_oicr0, iou_threshold, midn_weight, oicr_weight):
  """The whole of build_loss from the [B,P,ld] logits (needs num_stages <= 4 and stage k in columns
  col_oicr0 + k (C+1) ...).  class_logits / proba / scores0: the MIDN outputs (no gradient flows through them here; the
  MIDN backward is part of this node).  Returns (midn_loss, [oicr_loss_1 .. K], total, proposal_ind [K,B,C],
  proposal_labels [K,B,P,C+1], status)."""
  cfg = (int(num_classes), int(num_stages), int(col_r), int(col_c), int(col_oicr0), float(iou_threshold),
         float(midn_weight), float(oicr_weight))
  out = _LossHead.apply(logits_all, labels.contiguous(), num_proposals, proposals.contiguous(),
                        class_logits.detach().contiguous(), proba.detach().contiguous(), scores0.detach().contiguous(), cfg)
  K = cfg[1]
  return out[0], list(out[1:1 + K]), out[K + 1], out[K + 2], out[K + 3], out[K + 4]


# ---------------------------------------------------------------------------------------------
# K7: NMS post-process (core/builder.py:31-65)
# ---------------------------------------------------------------------------------------------
def multiclass_nms(boxes, scores, score_thresh, iou_thresh, max_size_per_class, max_total_size):
  """boxes [B,P,4], scores [B,P,C] -> (num_detections, boxes, scores, classes(1-based), index)."""
  require_cuda(boxes)
  boxes = boxes.contiguous()
  B, P, C = scores.shape
  sc_ptr, lds, _, _ = _rows(scores)
  dev = boxes.device
  lib = capi.load()
  nbytes = lib.c2d_nms_workspace_bytes(B, P, C, max_size_per_class)
  ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
  num = torch.empty((B,), dtype=torch.int32, device=dev)
  ob = torch.empty((B, max_total_size, 4), dtype=torch.float32, device=dev)
  osc = torch.empty((B, max_total_size), dtype=torch.float32, device=dev)
  oc = torch.empty((B, max_total_size), dtype=torch.float32, device=dev)
  oi = torch.empty((B, max_total_size), dtype=torch.int32, device=dev)
  call('c2d_multiclass_nms', ptr(boxes), sc_ptr, lds, B, P, C, float(score_thresh), float(iou_thresh),
       int(max_size_per_class), int(max_total_size), ptr(num), ptr(ob), ptr(osc), ptr(oc), ptr(oi), ptr(ws), nbytes,
       stream())
  return num, ob, osc, oc, oi


# ---------------------------------------------------------------------------------------------
# K8/K9: label kernels
# ---------------------------------------------------------------------------------------------
def label_lut(token_ids, lut, num_classes):
  require_cuda(token_ids, lut)
  B, T = token_ids.shape
  labels = torch.empty((B, num_classes), dtype=torch.float32, device=token_ids.device)
  call('c2d_label_lut', ptr(token_ids), B, T, ptr(lut), lut.numel(), num_classes, ptr(labels), stream())
  return labels


def wordvec_match(token_ids, emb, class_ids, exact_lut, return_similarity=False):
  require_cuda(token_ids, emb, class_ids, exact_lut)
  B, T = token_ids.shape
  V = emb.shape[0] - 1
  D = emb.shape[1]
  C = class_ids.numel()
  labels = torch.empty((B, C), dtype=torch.float32, device=token_ids.device)
  sim = torch.empty((B, C), dtype=torch.float32, device=token_ids.device) if return_similarity else None
  ws = torch.empty((max(1, capi.load().c2d_wordvec_workspace_bytes(B, T, C)),), dtype=torch.uint8, device=token_ids.device)
  call('c2d_wordvec_match', ptr(token_ids), B, T, ptr(emb), V, D, ptr(class_ids), C, ptr(exact_lut), ptr(labels),
       ptr(sim), ptr(ws), stream())
  return (labels, sim) if return_similarity else labels


def text_classifier_match(token_ids, emb, w1, b1, w2, b2, threshold, exact_lut, return_probas=False):
  """models/label_extractor.py:363-472 on pre-tokenised ids; w1 [D,H], w2 [H,C] (TF [in,out] layout)."""
  require_cuda(token_ids, emb, w1, b1, w2, b2, exact_lut)
  B, T = token_ids.shape
  V, D = emb.shape[0] - 1, emb.shape[1]
  H, C = w1.shape[1], w2.shape[1]
  labels = torch.empty((B, C), dtype=torch.float32, device=token_ids.device)
  probas = torch.empty((B, C), dtype=torch.float32, device=token_ids.device) if return_probas else None
  call('c2d_text_classifier_match', ptr(token_ids), B, T, ptr(emb), V, D, ptr(w1), ptr(b1), H, ptr(w2), ptr(b2), C,
       float(threshold), ptr(exact_lut), ptr(labels), ptr(probas), stream())
  return (labels, probas) if return_probas else labels
