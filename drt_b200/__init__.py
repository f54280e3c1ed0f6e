"""drt_b200 -- Blackwell-native differentiable refraction tracer (the hot path of lvjiahui/DRT).

    import drt_b200.DiffRender as Render      # drop-in for the reference's `import DiffRender as Render`
    from drt_b200 import optix                # drop-in for the JIT-built `optix` plugin module

The compute path is hand-written sm_100a CUDA in drt_b200/csrc behind the C ABI of
include/drt_b200.h; there is no CPU/PyTorch fallback.
"""
__version__ = "0.1.0"
