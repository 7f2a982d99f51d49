"""mapad_b200: B200-native hot path of mapAD behind a C ABI."""
import os as _os

# Several handles (one stream each) keep chunks in flight.  CUDA multiplexes streams onto 8 hardware queues by default;
# streams that share a queue falsely wait for each other's multi-second search kernels, so ask for the maximum (32)
# before the CUDA context is created.  Has no effect if the process initialised CUDA earlier.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
