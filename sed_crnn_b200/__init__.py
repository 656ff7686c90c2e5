"""sed_crnn_b200 -- B200-native (sm_100a) implementation of the sed-crnn hot path.

    feature          log-mel front end            (drop-in for reference feature.py:_mbe)
    crnn_lightning   TimePooledCRNN / FocalBCELoss / CRNNLightning   (reference crnn_lightning.py)
    sed              TimePooledCRNN / run_epoch   (reference sed.py)
    metrics          segment ER / F1              (reference metrics.py)
    engine           CRNNEngine: flat-buffer fused training step (+ NCCL data parallel)
    config           CRNNConfig and the BASELINE presets

Everything computes in libsedb200.so (include/sedb200.h); importing the package does not need a GPU,
running it does.
"""
__version__ = "0.1.0"
