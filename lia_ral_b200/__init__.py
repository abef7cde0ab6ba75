"""B200-native engine for LIA_RAL's GMM / i-vector / PLDA numeric hot path.

`lia_ral_b200.capi` binds the C ABI (include/lia_ral_b200.h) of liblia_ral_b200.so, the
hand-written sm_100a CUDA library built from csrc/.  There is no CPU implementation in this
package: the fp64 CPU restatement of the reference lives under oracle/ and is test
infrastructure only.
"""
__version__ = "0.1"
