"""realcamnet_b200 -- B200-native (sm_100a) implementation of RealCamNet's RAW->sRGB->bitstream
forward path behind the reference's own nn.Module API.

Modules mirror the reference files they stand in for:
  realcamnet_b200.raw2bit   <- models/raw2bit.py   (raw_compression_tcm_final, ConvTransBlock_mzj, ...)
  realcamnet_b200.tcm       <- models/tcm.py       (WMSA, Block, ConvTransBlock, SWAtten, ...)
  realcamnet_b200.LiteISP   <- models/LiteISP.py   (LiteISPNet_GFM_LSC, Res_GFM, ...)
  realcamnet_b200.groupmix  <- models/groupmix.py  (GMA_Block, EfficientAtt, ...)
  realcamnet_b200.networks  <- models/networks.py  (conv factory, RCAGroup, DWT)
  realcamnet_b200.layers / entropy_models <- the CompressAI subset the reference imports
All arithmetic runs in librcn_b200.so (C ABI: include/rcn_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
