"""computervision_codes_b200 -- B200-native (sm_100a) temporal head for CIAM-Group/ComputerVision_Codes.

Drop-in mirrors of the reference's temporal nn.Modules backed by hand-written CUDA kernels behind a
C ABI (include/tcn_b200.h).  Sub-packages:
  tcn    -- MT4MTLKD/Temporal_tenco/network.py and TERL/0_5fold_TCN_black/network.py
  kd     -- multi-teacher KD losses of MT4MTLKD/Spatial_cnn/run.py
  mstct  -- MT4MTLKD/Temporal_mstct/MSTCT/*.py and Temporal_mstct/network.py
"""
__version__ = "0.1.0"
