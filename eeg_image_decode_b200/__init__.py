"""B200-native (sm_100a) implementation of the EEG->image contrastive hot path of
dongyangli-del/EEG_Image_decode: ATM-S encoder + InfoNCE (ClipLoss) training step and retrieval eval."""
__version__ = "0.1.0"
