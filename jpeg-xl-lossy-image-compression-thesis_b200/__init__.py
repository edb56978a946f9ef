"""jxlb200 — B200-native JPEG XL lossy (VarDCT) encode hot path.

Host-side mirror of the one encoder call the thesis harness makes,
``DockerManager::execute_cjxl(input_file, output_file, distance, effort)``
(benchmark-jpegxl/src/docker_manager.rs:100-137), over the C ABI in
``include/jxlb200.h`` (``libjxlb200.so``).  There is no CPU fallback: importing
works anywhere, but creating an encoder without an sm_100 GPU raises.
"""
import os as _os

# The batch entry points keep up to 32 images in flight on separate CUDA streams; with the default of 8
# hardware work queues several streams share a queue and serialise (measured: 4-way instead of 8-way
# overlap).  Must be set before the CUDA context exists, hence at import time.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .encoder import (  # noqa: F401
    Encoder, EncodeError, Stats, PROPOSAL_NONE, PROPOSAL_PARTITIONING, PROPOSAL_FACTORED_ENTROPY,
    PROPOSAL_COMBINED, FLAG_FIXED_DCT8, FLAG_UNIFORM_QF, FLAG_QUALITY, FLAG_FORCED_ACS, FLAG_GABORISH, FLAG_CFL, STAGES, frame_dims, library_path, load_library,
)
from .synth import synth_image, synth_batch  # noqa: F401
from .sharding import (  # noqa: F401
    shard_indices, distance_for_image, ShardStats, JobStats, gather_stats, bind_to_gpu_numa_node,
)
