"""libllsm2_b200: B200-native (sm_100a CUDA) implementation of libllsm2's layer-0 analysis /
synthesis hot path behind the reference's C API. See DESIGN.md and include/llsm_b200.h.

Python side = thin host mirror over the C ABI (ctypes); torch is used only for device memory,
streams and torch.distributed."""
from .api import (Context, synthesize_l0, synthesize_l0_host, synthesize_harmonics, output_length,  # noqa: F401
                  analyze_l0, analyze_l0_host, anasynth_host, tolayer1, tolayer0, synthesize_l1, synthesize_l0_shard, RtSynth,
                  chunk_phasepropagate, chunk_phasesync_rps, coder_encode, coder_decode, frames_to_blob, blob_to_frames,
                  frames_stretch, stretch_map)
from ._lib import LlsmB200Error  # noqa: F401
