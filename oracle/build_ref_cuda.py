"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Compiles the reference's own CUDA kernels (unmodified, from where they lie under /root/reference) for sm_100a
into oracle/_ref/fpc_ref_ransac_voting.so with our tiny binding (oracle/ref_binding.cpp).  Only possible in the
build container; the built module travels to the GPU box (oracle/_ref is git-ignored, not gpurun-ignored), where
tests/test_reference_cuda_gpu.py checks the product's FPC_ARITH_NVCC_FMA mode against it bit for bit.
Never uses the reference's build system; never copies reference sources into the repo."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CU = "/root/reference/source_code/FastPoseCNN/lib/ransac_voting_gpu_layer/src/ransac_voting_kernel.cu"
OUT = os.path.join(HERE, "_ref")
NAME = "fpc_ref_ransac_voting"


def available() -> bool:
    return os.path.exists(os.path.join(OUT, NAME + ".so"))


def load():
    """Imports the built module (GPU box or build container).  Returns None if it was never built."""
    if not available():
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(OUT, NAME + ".so"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build(verbose: bool = False) -> bool:
    if not os.path.exists(REF_CU):
        return False
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ["CC"] = "/usr/bin/gcc"
    os.environ["CXX"] = "/usr/bin/g++"
    from torch.utils.cpp_extension import load as jit_load
    jit_load(name=NAME, sources=[os.path.join(HERE, "ref_binding.cpp"), REF_CU], build_directory=OUT,
             extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++"],
             extra_include_paths=[os.path.dirname(REF_CU)], verbose=verbose, is_python_module=False)
    return available()


if __name__ == "__main__":
    ok = build(verbose="-v" in sys.argv)
    print("built" if ok else "reference sources not available: nothing built")
