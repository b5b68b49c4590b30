"""transmf_ad_b200 -- B200-native (sm_100a) implementation of the TransMF_AD training hot path.

* ``transmf_ad_b200.models``      drop-in mirror of the reference ``models`` package (nn.Module API unchanged)
* ``transmf_ad_b200.functional``  autograd Functions over the C ABI (``include/tmf.h`` -> ``libtmf_sm100a.so``)
* ``transmf_ad_b200.dp``          data-parallel gradient bucketing over NCCL
* ``transmf_ad_b200.csrc``        hand-written CUDA kernels
There is no CPU / library fallback: ops raise if the CUDA library is missing or the device is not cc 10.x.
"""
__version__ = "0.1.0"
