"""distgcn_b200: B200-native GCN-scored local-greedy MWIS path of zhongyuanzhao/distgcn.

Layout (only what the hot path needs):
  csrc/               CUDA kernels + the C-ABI (include/distgcn_b200.h), built to libdistgcn_b200.so
  _lib / engine       ctypes binding and the thin object layer over it
  batch / ckpt        packed-CSR ingest, TensorFlow-bundle checkpoint reader (no TensorFlow)
  layers / models / heuristics / mwis_dqn_call / mwis_gdpg_call / runtime_config / directory
                      host-side mirrors of the reference's operator API for this path
  shard               graph-batch sharding across GPUs (no collectives)
"""
__version__ = "0.1.0"
