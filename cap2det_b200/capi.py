"""ctypes binding of the C ABI in include/cap2det_b200.h.

There is NO CPU fallback: if the shared library cannot be built/loaded, or a call
returns a non-zero status, a RuntimeError / ValueError is raised.
"""
import ctypes
import os

import torch

from cap2det_b200 import build as _build

C2D_F32, C2D_BF16, C2D_U8 = 0, 1, 2
MASKED_MAX, MASKED_MIN, MASKED_SUM, MASKED_AVG, MASKED_ARGMAX, MASKED_ARGMIN = range(6)

_c_int, _c_float, _c_ll, _c_sz, _p = (ctypes.c_int, ctypes.c_float, ctypes.c_longlong,
                                      ctypes.c_size_t, ctypes.c_void_p)

# name -> (restype, argtypes); mirrors include/cap2det_b200.h one to one.
SIGNATURES = {
    'c2d_version': (_c_int, []),
    'c2d_last_error': (ctypes.c_char_p, []),
    'c2d_has_tensor_core_head': (_c_int, []),
    'c2d_profile_enable': (None, [_c_int]),
    'c2d_profile_reset': (None, []),
    'c2d_profile_read': (_c_int, [_c_int, _p, _p, _p]),
    'c2d_profile_entry': (_c_int, [_c_int, _p, _p, _p]),
    'c2d_launch_count': (_c_ll, []),
    'c2d_reset_launch_count': (None, []),
    'c2d_box_area': (_c_int, [_p, _c_int, _p, _p]),
    'c2d_box_intersect': (_c_int, [_p, _p, _c_int, _p, _p]),
    'c2d_box_iou': (_c_int, [_p, _p, _c_int, _p, _p]),
    'c2d_box_flip_left_right': (_c_int, [_p, _c_int, _p, _p]),
    'c2d_box_scale_to_new_size': (_c_int, [_p, _c_int, _c_int, _c_int, _c_int, _c_int, _p, _p]),
    'c2d_masked_reduce': (_c_int, [_p, _p, _c_int, _c_int, _c_int, _c_int, _p, _p, _p]),
    'c2d_masked_max_bwd': (_c_int, [_p, _p, _c_int, _c_int, _c_int, _p, _p, _p]),
    'c2d_masked_softmax': (_c_int, [_p, _p, _c_int, _c_int, _c_int, _p, _p]),
    'c2d_roi_crop_maxpool_fwd': (_c_int, [_p, _c_int, _c_int, _c_int, _c_int, _p, _c_int, _c_int, _c_int,
                                          _c_int, _p, _c_int, _p]),
    'c2d_roi_crop_maxpool_bwd': (_c_int, [_p, _c_int, _c_int, _c_int, _c_int, _p, _c_int, _c_int, _c_int,
                                          _c_int, _p, _c_int, _p, _p]),
    'c2d_roi_argmax_code_bytes': (_c_sz, [_c_int, _c_int, _c_int]),
    'c2d_roi_crop_maxpool_fwd_codes': (_c_int, [_p, _c_int, _c_int, _c_int, _c_int, _p, _c_int, _c_int, _c_int, _c_int, _p, _c_int, _p, _p]),
    'c2d_roi_crop_maxpool_bwd_codes': (_c_int, [_c_int, _c_int, _c_int, _c_int, _p, _c_int, _c_int, _c_int, _c_int, _p, _p, _c_int, _p, _p]),
    'c2d_head_num_convs': (_c_int, []),
    'c2d_head_conv_spec': (_c_int, [_c_int, _p, _p, _p, _p, _p]),
    'c2d_head_param_floats': (_c_ll, []),
    'c2d_head_param_offsets': (_c_int, [_c_int, _p, _p, _p, _p, _p]),
    'c2d_head_workspace_bytes': (_c_sz, [_c_int, _c_int]),
    'c2d_head_mixed5_fwd': (_c_int, [_p, _c_int, _c_int, _p, _p, _c_sz, _p, _c_float, _p, _p]),
    'c2d_head_mixed5_bwd': (_c_int, [_p, _c_int, _c_int, _p, _p, _c_sz, _p, _c_float, _p, _p, _p, _p]),
    'c2d_head_mixed5_bwd_fold': (_c_int, [_p, _c_int, _c_int, _p, _p, _c_sz, _p, _c_float, _p, _p, _p, _p, _p, _p, _p]),
    'c2d_roi_crop_maxpool_bwd_codes_fold': (_c_int, [_c_int, _c_int, _c_int, _c_int, _p, _c_int, _c_int, _c_int, _c_int, _p, _p, _p, _p, _c_int, _p, _p]),
    'c2d_roi_bwd_tiles_workspace_bytes': (_c_sz, [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int]),
    'c2d_roi_crop_maxpool_bwd_tiles': (_c_int, [_c_int, _c_int, _c_int, _c_int, _p, _c_int, _c_int, _c_int, _c_int, _p, _p, _c_int,
                                                _p, _p, _c_int, _p, _c_sz, _p, _p]),
    'c2d_resize_bilinear': (_c_int, [_p, _c_int, _c_int, _c_int, _c_int, _c_int, _p, _c_int, _c_int, _p]),
    'c2d_image_flip_left_right': (_c_int, [_p, _c_int, _c_int, _c_int, _c_int, _c_int, _p, _p, _p]),
    'c2d_box_scale_batch': (_c_int, [_p, _p, _c_int, _c_int, _c_int, _c_int, _p, _p]),
    'c2d_backbone_num_convs': (_c_int, []),
    'c2d_backbone_conv_spec': (_c_int, [_c_int, _p, _p, _p, _p, _p]),
    'c2d_backbone_param_floats': (_c_ll, []),
    'c2d_backbone_param_offsets': (_c_int, [_c_int, _p, _p, _p, _p, _p]),
    'c2d_backbone_out_dims': (_c_int, [_c_int, _c_int, _p, _p]),
    'c2d_backbone_workspace_bytes': (_c_sz, [_c_int, _c_int, _c_int]),
    'c2d_backbone_fwd': (_c_int, [_p, _c_int, _c_int, _c_int, _p, _p, _c_sz, _p, _p]),
    'c2d_backbone_bwd': (_c_int, [_p, _p, _c_int, _c_int, _c_int, _p, _p, _c_sz, _p, _p]),
    'c2d_backbone_mixed4e_input': (_c_int, [_p, _c_int, _c_int, _c_int, _p, _p]),
    'c2d_conv_img_bf16_fwd': (_c_int, [_p, _c_int, _c_int, _c_int, _c_int, _c_int, _p, _c_int, _c_int, _c_int, _p, _c_int, _p, _c_int, _p]),
    'c2d_conv_img_bf16_dgrad': (_c_int, [_p, _c_int, _c_int, _c_int, _c_int, _c_int, _p, _c_int, _p, _p, _c_int, _p]),
    'c2d_conv_img_bf16_wgrad': (_c_int, [_p, _c_int, _p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _p, _p, _p]),
    'c2d_conv_bf16_fwd': (_c_int, [_p, _c_int, _c_int, _c_int, _c_int, _p, _c_int, _c_int, _c_int, _p, _c_int, _p, _c_int, _p]),
    'c2d_conv_bf16_dgrad': (_c_int, [_p, _c_int, _c_int, _c_int, _c_int, _p, _c_int, _c_int, _c_int, _p, _c_int, _c_int, _p]),
    'c2d_conv_bf16_wgrad': (_c_int, [_p, _c_int, _p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _p, _p]),
    'c2d_fc_workspace_bytes': (_c_sz, [_c_int, _c_int, _c_int, _c_int]),
    'c2d_fc_fwd': (_c_int, [_p, _c_int, _c_int, _p, _p, _c_int, _p, _c_int, _c_int, _p, _c_sz, _p]),
    'c2d_fc_bwd': (_c_int, [_p, _c_int, _c_int, _p, _c_int, _p, _c_int, _p, _p, _p, _c_int, _p, _c_sz, _p]),
    'c2d_midn_fwd': (_c_int, [_p, _p, _c_int, _p, _c_int, _c_int, _c_int, _p, _p, _p, _p]),
    'c2d_midn_bwd': (_c_int, [_p, _c_int, _p, _c_int, _c_int, _c_int, _p, _p, _p, _p, _p, _p, _p, _c_int, _p]),
    'c2d_sigmoid_ce_mean_fwd': (_c_int, [_p, _p, _c_int, _c_float, _p, _p]),
    'c2d_sigmoid_ce_mean_bwd': (_c_int, [_p, _p, _c_int, _c_float, _p, _p, _p]),
    'c2d_softmax_rows': (_c_int, [_p, _c_int, _c_int, _c_int, _p, _c_int, _p]),
    'c2d_oicr_assign': (_c_int, [_p, _p, _p, _p, _c_int, _c_float, _c_int, _c_int, _c_int, _p, _p, _p, _p]),
    'c2d_oicr_ce_fwd': (_c_int, [_p, _p, _c_int, _p, _c_int, _c_int, _c_int, _c_float, _p, _p]),
    'c2d_oicr_ce_bwd': (_c_int, [_p, _p, _c_int, _p, _c_int, _c_int, _c_int, _c_float, _p, _p, _c_int, _p]),
    'c2d_nms_workspace_bytes': (_c_sz, [_c_int, _c_int, _c_int, _c_int]),
    'c2d_multiclass_nms': (_c_int, [_p, _p, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float, _c_int, _c_int,
                                    _p, _p, _p, _p, _p, _p, _c_sz, _p]),
    'c2d_label_lut': (_c_int, [_p, _c_int, _c_int, _p, _c_int, _c_int, _p, _p]),
    'c2d_text_classifier_match': (_c_int, [_p, _c_int, _c_int, _p, _c_int, _c_int, _p, _p, _c_int, _p, _p, _c_int,
                                           _c_float, _p, _p, _p, _p]),
    'c2d_optimizer_update': (_c_int, [_c_int, _p, _p, _p, _p, _p, _c_ll, _c_float, _c_float, _c_float, _c_float, _c_float,
                                      _c_float, _c_int, _p]),
    'c2d_ema_update': (_c_int, [_p, _p, _c_ll, _c_float, _p]),
    'c2d_dropout_apply': (_c_int, [_p, _p, _c_float, _p, _c_ll, _p]),
    'c2d_adagrad_update': (_c_int, [_p, _p, _p, _c_ll, _c_float, _c_float, _c_float, _p]),
    'c2d_l2_loss': (_c_int, [_p, _c_ll, _c_float, _p, _p]),
    'c2d_dropout_keep_mask': (_c_int, [_p, ctypes.c_uint, _c_ll, _c_float, _p, _p]),
    'c2d_l2_loss_add': (_c_int, [_p, _c_ll, _c_float, _p, _p, _p]),
    'c2d_loss_head_fwd': (_c_int, [_p, _c_int, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float,
                                   _c_float, _p, _p, _p, _p, _p, _p]),
    'c2d_loss_head_bwd': (_c_int, [_p, _c_int, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                   _c_float, _c_float, _p, _p, _p, _p, _p, _p, _p, _p]),
    'c2d_wordvec_workspace_bytes': (_c_sz, [_c_int, _c_int, _c_int]),
    'c2d_wordvec_match': (_c_int, [_p, _c_int, _c_int, _p, _c_int, _c_int, _p, _c_int, _p, _p, _p, _p, _p]),
}

_lib = None


def library_path():
  return _build.LIB_PATH


def load(build_if_missing=True):
  """Loads (building in-tree first when nvcc is present and sources changed) the CUDA library."""
  global _lib
  if _lib is not None:
    return _lib
  path = _build.LIB_PATH
  if build_if_missing and os.path.exists(_build.NVCC):
    path = _build.build_library()
  if not os.path.exists(path):
    raise RuntimeError(
        'cap2det_b200: CUDA library %s is missing and cannot be built (no nvcc); '
        'there is no CPU fallback.' % path)
  lib = ctypes.CDLL(path)
  for name, (res, args) in SIGNATURES.items():
    fn = getattr(lib, name)          # AttributeError if the export is missing
    fn.restype = res
    fn.argtypes = args
  _lib = lib
  return lib


class C2DError(RuntimeError):
  pass


def check(status):
  if status == 0:
    return
  msg = load().c2d_last_error().decode('utf-8', 'replace')
  if status == -1:
    raise ValueError('cap2det_b200: ' + msg)
  raise C2DError('cap2det_b200 (status %d): %s' % (status, msg))


def ptr(t):
  """Device pointer of a tensor (None -> NULL)."""
  if t is None:
    return None
  return ctypes.c_void_p(t.data_ptr())


def stream():
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
  check(getattr(load(), name)(*args))


def require_cuda(*tensors):
  """All tensors of one call must be contiguous, on ONE CUDA device, and that device must be the current one: the
  library launches on the current device / current stream and takes raw pointers, so a tensor of another GPU would
  be dereferenced on the wrong device.  (Use `with torch.cuda.device(t.device):` or torch.cuda.set_device.)"""
  device = None
  for t in tensors:
    if t is None:
      continue
    if not t.is_cuda:
      raise RuntimeError('cap2det_b200: tensors must live on a CUDA device (no CPU fallback)')
    if not t.is_contiguous():
      raise ValueError('cap2det_b200: tensors must be contiguous')
    if device is None:
      device = t.device
    elif t.device != device:
      raise ValueError('cap2det_b200: tensors of one call live on different devices (%s and %s)' % (device, t.device))
  if device is not None and device.index != torch.cuda.current_device():
    raise RuntimeError('cap2det_b200: tensors live on %s but the current CUDA device is cuda:%d; wrap the call in '
                       'torch.cuda.device(%r)' % (device, torch.cuda.current_device(), str(device)))


def dtype_code(dtype):
  if dtype == torch.float32:
    return C2D_F32
  if dtype == torch.bfloat16:
    return C2D_BF16
  if dtype == torch.uint8:
    return C2D_U8
  raise ValueError('unsupported dtype %s' % dtype)


def launch_count():
  return int(load().c2d_launch_count())


def reset_launch_count():
  load().c2d_reset_launch_count()
