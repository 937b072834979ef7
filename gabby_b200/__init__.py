"""gabby_b200 -- B200-native (sm_100a) Llama-3 forward behind gabby's inference::Generator.

The product is native: gabby_b200/csrc (CUDA kernels + the b2l C-ABI, include/b2l.h) and
gabby_b200/host (C++ host layer mirroring /root/reference/src/inference/generator.h). This
Python package only carries the build recipe, ctypes bindings for tests/bench, and the
synthetic-model writer. There is no CPU fallback; see gabby_b200._capi.lib().
"""
__all__ = ["synth", "build", "_capi"]
