"""Builds libdrt_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
SO = os.path.join(OUT_DIR, "libdrt_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]


def sources():
    return sorted(os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cu", ".cuh"))) + [
        os.path.join(os.path.dirname(HERE), "include", "drt_b200.h")]


def source_hash():
    """sha1 over the kernel sources and the C header: profiles/ncu_summary.json records it, bench.py refuses ncu-derived
    figures (roofline.traffic, lane statistics) that were captured from different kernels."""
    import hashlib
    h = hashlib.sha1()
    for f in sources():
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(os.path.basename(f).encode())
            h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, os.path.join(SRC, "capi.cu")]
    env = dict(os.environ)
    env.pop("CC", None)  # the image exports CC=/opt/gcc/bin/gcc; let nvcc pick the system g++
    env.pop("CXX", None)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libdrt_b200.so")
    return SO


def load_torch_plugin():
    """Imports the in-tree compiled plugin (_C/optix_b200/optix_b200.so), building it first when absent or stale."""
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    so = os.path.join(OUT_DIR, "optix_b200", "optix_b200.so")
    src = os.path.join(SRC, "optix_extend_b200.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src) or needs_build():
        return build_torch_plugin()
    spec = importlib.util.spec_from_file_location("optix_b200", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build_torch_plugin(verbose=False):
    """The reference's pybind plugin class (optix_extend.cpp:77-83) as a compiled torch extension over the C ABI
    (csrc/optix_extend_b200.cpp, the binding of INTEGRATION.md section 4): built IN-TREE into _C/optix_b200/ so that it
    travels to the GPU box; returns the imported module (`module.optix_mesh`)."""
    import torch.utils.cpp_extension as ce
    build()
    out = os.path.join(OUT_DIR, "optix_b200")
    os.makedirs(out, exist_ok=True)
    env_cc = {k: os.environ.pop(k) for k in ("CC", "CXX") if k in os.environ}
    try:
        return ce.load(name="optix_b200", sources=[os.path.join(SRC, "optix_extend_b200.cpp")],
                       extra_include_paths=[os.path.join(os.path.dirname(HERE), "include")],
                       extra_ldflags=[f"-L{OUT_DIR}", "-ldrt_b200", f"-Wl,-rpath,{OUT_DIR}", "-Wl,-rpath,$ORIGIN/.."],
                       build_directory=out, with_cuda=True, verbose=verbose)
    finally:
        os.environ.update(env_cc)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
