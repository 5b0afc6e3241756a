"""Host-side mirror of the reference's native-op interface for the Box Seeker path.

    reference module                                          this package
    pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils      ->   pcdet_ops.roiaware_pool3d_utils
    pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda       ->   pcdet_ops.roiaware_pool3d_cuda
    pcdet.ops.iou3d_nms.iou3d_nms_utils                  ->   pcdet_ops.iou3d_nms_utils
    pcdet.ops.iou3d_nms.iou3d_nms_cuda                   ->   pcdet_ops.iou3d_nms_cuda

Same function names, argument meaning and return types; everything executes in
libfnp_sm100.so (sm_100a) through the C ABI of include/fnp.h.  See INTEGRATION.md.
"""
from . import iou3d_nms_cuda, iou3d_nms_utils, roiaware_pool3d_cuda, roiaware_pool3d_utils  # noqa: F401
