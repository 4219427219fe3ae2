"""hosnerf_b200 - B200-native (sm_100a) kernels for the HOSNeRF per-ray render hot path,
behind the reference's own module surface (see DESIGN.md / INTEGRATION.md).

    from hosnerf_b200 import MipNeRF360, LitMipNeRF360, Network

The compute lives in ``libhosnerf_b200.so`` (C ABI: include/hosnerf_b200.h); importing this
package does not need the library, calling any op does - there is no CPU fallback.
"""
from .mip360 import (LitMipNeRF360, MipNeRF360, MipNeRF360MLP, NeRFMLP, PropMLP, select_state_index,
                     set_precision)
from .human import (BodyPoseRefiner, CanonicalMLP, MotionBasisComputer, MotionWeightVolumeDecoder, Network,
                    NonRigidForwardMLP, NonRigidMotionMLP, default_cfg)

from .hosnerf import cycle_loss, flow_loss, render_hosnerf_chunk, train_hosnerf_chunk

__all__ = ["render_hosnerf_chunk", "train_hosnerf_chunk", "flow_loss", "cycle_loss", "LitMipNeRF360", "MipNeRF360", "MipNeRF360MLP", "NeRFMLP", "PropMLP", "Network", "CanonicalMLP",
           "NonRigidMotionMLP", "NonRigidForwardMLP", "MotionWeightVolumeDecoder", "BodyPoseRefiner",
           "MotionBasisComputer", "default_cfg", "select_state_index", "set_precision"]
