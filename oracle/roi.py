"""Oracle (test infrastructure): ROI feature extraction = crop_and_resize + max_pool.

Restates ``tf.image.crop_and_resize`` (bilinear, extrapolation_value 0) as called at
``models/utils.py:147-155`` and ``slim.max_pool2d`` (VALID) at ``models/utils.py:157-160``.
TensorFlow is not under /root/reference and cannot be installed here, so this is
**parity unpinned**; it follows the TF 1.15 CPU kernel (crop_and_resize_op.cc)
expression by expression, every fp32 op rounded separately (the CPU build does not
contract FMAs):

  hs   = (y2 - y1) * (H - 1) / (ch - 1)              (ch > 1)
  in_y = y1 * (H - 1) + y * hs
  row is all-zero if in_y < 0 or in_y > H - 1
  top = floor(in_y), bot = ceil(in_y), yl = in_y - top        (same for x)
  t = TL + (TR - TL) * xl ; b = BL + (BR - BL) * xl ; out = t + (b - t) * yl

Max-pool backward routes to the FIRST maximum in row-major window order.
crop_and_resize backward is accumulated in float64 in ascending (n, y, x) order
(TF's GPU kernel uses atomics => order undefined; compared at 1e-5).
"""
import numpy as np

F = np.float32


def _sample_coords(lo, hi, size, n_out):
  """Returns (in_coord [N,n_out] f32, valid [N,n_out] bool)."""
  lo = lo.astype(np.float32)
  hi = hi.astype(np.float32)
  sm1 = F(size - 1)
  if n_out > 1:
    scale = (hi - lo) * sm1 / F(n_out - 1)
    idx = np.arange(n_out, dtype=np.float32)[None, :]
    coord = (lo * sm1)[:, None] + idx * scale[:, None]
  else:
    coord = (F(0.5) * (lo + hi) * sm1)[:, None]
  valid = ~((coord < F(0)) | (coord > sm1))
  return coord.astype(np.float32), valid


def crop_and_resize(fmap, boxes, box_ind, crop_size):
  """fmap [B,H,W,C] f32, boxes [N,4] (y1,x1,y2,x2) normalised, box_ind [N] -> [N,ch,cw,C]."""
  fmap = np.asarray(fmap, np.float32)
  boxes = np.asarray(boxes, np.float32)
  B, H, W, C = fmap.shape
  ch, cw = crop_size
  N = boxes.shape[0]
  in_y, vy = _sample_coords(boxes[:, 0], boxes[:, 2], H, ch)
  in_x, vx = _sample_coords(boxes[:, 1], boxes[:, 3], W, cw)
  in_y = np.where(vy, in_y, F(0)); in_x = np.where(vx, in_x, F(0))
  top = np.floor(in_y).astype(np.int64); bot = np.ceil(in_y).astype(np.int64)
  left = np.floor(in_x).astype(np.int64); right = np.ceil(in_x).astype(np.int64)
  yl = (in_y - top.astype(np.float32)).astype(np.float32)[:, :, None, None]
  xl = (in_x - left.astype(np.float32)).astype(np.float32)[:, None, :, None]
  bi = np.asarray(box_ind, np.int64)[:, None, None]
  TL = fmap[bi, top[:, :, None], left[:, None, :]]
  TR = fmap[bi, top[:, :, None], right[:, None, :]]
  BL = fmap[bi, bot[:, :, None], left[:, None, :]]
  BR = fmap[bi, bot[:, :, None], right[:, None, :]]
  t = TL + (TR - TL) * xl
  b = BL + (BR - BL) * xl
  out = t + (b - t) * yl
  valid = (vy[:, :, None] & vx[:, None, :])[..., None]
  return np.where(valid, out, F(0)).astype(np.float32)


def max_pool_2x2(x):
  """slim.max_pool2d(k=2, s=2, VALID): [N,2h,2w,C] -> [N,h,w,C]."""
  N, Hc, Wc, C = x.shape
  h, w = Hc // 2, Wc // 2
  x = x[:, :2 * h, :2 * w].reshape(N, h, 2, w, 2, C)
  return x.max(axis=(2, 4))


def roi_crop_maxpool_fwd(fmap, proposals, crop=14, chunk=128):
  """models/utils.py:147-160. proposals [B,P,4] -> [B*P, crop/2, crop/2, C]."""
  fmap = np.asarray(fmap, np.float32)
  B, P, _ = proposals.shape
  boxes = np.asarray(proposals, np.float32).reshape(-1, 4)
  box_ind = np.repeat(np.arange(B), P)   # models/utils.py:148-149
  outs = []
  for s in range(0, boxes.shape[0], chunk):
    c = crop_and_resize(fmap, boxes[s:s + chunk], box_ind[s:s + chunk], (crop, crop))
    outs.append(max_pool_2x2(c))
  return np.concatenate(outs, axis=0)


def roi_crop_maxpool_bwd(fmap, proposals, dout, crop=14, chunk=64):
  """Gradient of roi_crop_maxpool_fwd w.r.t. fmap.  dout [B*P, crop/2, crop/2, C]."""
  fmap = np.asarray(fmap, np.float32)
  B, H, W, C = fmap.shape
  P = proposals.shape[1]
  boxes = np.asarray(proposals, np.float32).reshape(-1, 4)
  box_ind = np.repeat(np.arange(B), P)
  dfm = np.zeros((B, H, W, C), np.float64)
  hp = crop // 2
  cidx = np.arange(C)
  for s in range(0, boxes.shape[0], chunk):
    bx = boxes[s:s + chunk]; bi = box_ind[s:s + chunk]
    n = bx.shape[0]
    cr = crop_and_resize(fmap, bx, bi, (crop, crop))
    win = cr.reshape(n, hp, 2, hp, 2, C).transpose(0, 1, 3, 2, 4, 5).reshape(n, hp, hp, 4, C)
    arg = np.argmax(win, axis=3)     # first max in row-major (dy,dx) order
    g = np.zeros((n, hp, hp, 4, C), np.float32)
    np.put_along_axis(g, arg[:, :, :, None, :], np.asarray(dout[s:s + n], np.float32)[:, :, :, None, :], axis=3)
    dcrop = g.reshape(n, hp, hp, 2, 2, C).transpose(0, 1, 3, 2, 4, 5).reshape(n, crop, crop, C)
    in_y, vy = _sample_coords(bx[:, 0], bx[:, 2], H, crop)
    in_x, vx = _sample_coords(bx[:, 1], bx[:, 3], W, crop)
    in_y = np.where(vy, in_y, F(0)); in_x = np.where(vx, in_x, F(0))
    top = np.floor(in_y).astype(np.int64); bot = np.ceil(in_y).astype(np.int64)
    left = np.floor(in_x).astype(np.int64); right = np.ceil(in_x).astype(np.int64)
    yl = (in_y - top.astype(np.float32)).astype(np.float32)
    xl = (in_x - left.astype(np.float32)).astype(np.float32)
    valid = (vy[:, :, None] & vx[:, None, :])
    dcrop = np.where(valid[..., None], dcrop, F(0))
    w_tl = ((F(1) - yl)[:, :, None] * (F(1) - xl)[:, None, :])
    w_tr = ((F(1) - yl)[:, :, None] * xl[:, None, :])
    w_bl = (yl[:, :, None] * (F(1) - xl)[:, None, :])
    w_br = (yl[:, :, None] * xl[:, None, :])
    bi3 = np.broadcast_to(bi[:, None, None], (n, crop, crop))
    for (yy, xx, w) in ((top, left, w_tl), (top, right, w_tr), (bot, left, w_bl), (bot, right, w_br)):
      y3 = np.broadcast_to(yy[:, :, None], (n, crop, crop))
      x3 = np.broadcast_to(xx[:, None, :], (n, crop, crop))
      np.add.at(dfm, (bi3.reshape(-1), y3.reshape(-1), x3.reshape(-1)),
                (dcrop * w[..., None]).reshape(-1, C).astype(np.float64))
  del cidx
  return dfm.astype(np.float32)
