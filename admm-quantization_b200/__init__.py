"""B200-native quantization-aware CP factorization (hot path of KamikaziZen/admm-quantization).

Layout: `csrc/` CUDA kernels + C ABI (-> `lib/libadmmq.so`), `source/` the drop-in Python
modules (`source.admm`, `source.quantization`, `source.parafac_epc`, `source.utils`) and
`scripts/factorize.py`, the reference CLI.  Put this directory on `sys.path` and use
`from source.admm import admm_iteration` exactly as with the reference.
"""
