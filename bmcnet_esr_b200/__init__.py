"""bmcnet_esr_b200 -- B200 (sm_100a) implementation of the BMCNet inference hot path.

Drop-in mirrors of the reference's Python surface:
    from bmcnet_esr_b200.models.BMCNet import BMCNet              # models/BMCNet.py
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain  # models/BMCNet_plain.py
    from bmcnet_esr_b200.dataloader.encodings import *            # dataloader/encodings.py
Everything computes through libbmc_b200.so (include/bmc_b200.h); there is no CPU fallback.
"""
__version__ = '0.1.0'
