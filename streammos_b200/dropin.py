"""Install the sm_100a implementations under the names the reference imports, so that
models/StreamMOS.py, networks/multi_view_encoder.py and deformattn/ run unmodified:

    import streammos_b200.dropin as dropin; dropin.install()
    # then:  import deep_point; import MultiScaleDeformableAttention; from networks import backbone

(The reference has no plugin registry; its modules are found through sys.modules.)

`point_major_points` (default): BilinearSample and the fused PointNetStacker return their (B, C, N, 1) results with channels_last strides — same shape
and values, each point's channels contiguous, the layout the next VoxelMaxPool and the 1x1 point convolutions read
fastest. Pass False for (B, C, N, 1)-contiguous results as the reference produces them."""
import sys


def install(point_major_points=True):
    from . import MultiScaleDeformableAttention as msda
    from . import backbone as b200_backbone
    from . import deep_point as b200_deep_point
    from . import point_deep as b200_point_deep

    sys.modules["point_deep"] = b200_point_deep
    sys.modules["point_deep.cuda_kernel"] = b200_point_deep.cuda_kernel
    sys.modules["point_deep.cpu_kernel"] = b200_point_deep.cpu_kernel
    sys.modules["deep_point"] = b200_deep_point
    sys.modules["MultiScaleDeformableAttention"] = msda
    b200_backbone.BilinearSample.point_major_out = bool(point_major_points)
    b200_backbone.PointNetStacker.point_major_out = bool(point_major_points)
    # networks.backbone.BilinearSample is looked up by name at model build time (backbone.py:37-43,
    # multi_view_encoder.py:377-378): replace the class if the reference package is importable.
    ref_backbone = sys.modules.get("networks.backbone")
    if ref_backbone is None:
        try:
            import networks.backbone as ref_backbone  # noqa: F811
        except Exception:
            ref_backbone = None
    if ref_backbone is not None:
        ref_backbone.BilinearSample = b200_backbone.BilinearSample
        # models/StreamMOS.py:77 builds `backbone.PointNetStacker(7, C, pre_bn=True, stack_num=2)`: same constructor and
        # parameter names, eval forward fused into one kernel (training / autograd keep the torch layers)
        ref_backbone.PointNetStacker = b200_backbone.PointNetStacker
    # deformattn/modules/ms_deform_attn.py is pure Python around the compiled sampling core; its forward is swapped for
    # the one that fuses softmax + sampling-location arithmetic into the kernel at inference (modules.py). Patching the
    # method (not the class) also covers modules built before install() and `from deformattn.modules import MSDeformAttn`
    # bindings made at import time (networks/multi_view_encoder.py:8).
    try:
        import deformattn.modules.ms_deform_attn as ref_msda_mod
        from . import modules as b200_modules
        ref_msda_mod.MSDeformAttn.forward = b200_modules.msdeformattn_forward
    except Exception:
        pass
    return True
