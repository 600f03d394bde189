"""rl_ctr_prediction_b200 -- the B200 (sm_100a) hot path of jqsl2012/RL_CTR_Prediction behind the
reference's own Python surface (model classes, optimizer hand-off, loop functions).

    from rl_ctr_prediction_b200 import p_model, optim, pretrain_main

Everything that computes is a call into ``librlctr_sm100a.so`` (``include/rlctr.h``); there is no CPU
or eager-PyTorch fallback.  Build the library with ``python -m rl_ctr_prediction_b200.build``.
"""
__version__ = "0.1.0"
