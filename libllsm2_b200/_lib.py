"""Loader of the CUDA shared library. There is deliberately no fallback: if libllsm2_b200.so is
missing or no CUDA device is present the package raises."""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libllsm2_b200.so")
_lib = None


class LlsmB200Error(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise LlsmB200Error(
            "libllsm2_b200.so is not built (run libllsm2_b200/build.sh or __graft_entry__.build()); "
            "this package has no CPU path")
    L = C.CDLL(_SO)
    L.llsm_b200_create.restype = C.c_void_p
    L.llsm_b200_create.argtypes = [C.c_int]
    L.llsm_b200_destroy.argtypes = [C.c_void_p]
    L.llsm_b200_last_error.restype = C.c_char_p
    L.llsm_b200_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.llsm_b200_synchronize.argtypes = [C.c_void_p]
    L.llsm_b200_launch_count.restype = C.c_longlong
    L.llsm_b200_launch_count.argtypes = [C.c_void_p]
    L.llsm_b200_output_length.argtypes = [C.c_int, C.c_float, C.c_float]
    L.llsm_b200_template_length.argtypes = [C.c_int]
    P = C.c_void_p
    L.llsm_b200_synthesize_l0.argtypes = [P, C.POINTER(abi.Conf), C.POINTER(abi.Frames),
                                          C.POINTER(abi.SOptions), C.POINTER(abi.Output)]
    L.llsm_b200_synthesize_l0_host.argtypes = L.llsm_b200_synthesize_l0.argtypes
    L.llsm_b200_synthesize_l0_shard.argtypes = L.llsm_b200_synthesize_l0.argtypes + [C.c_int, C.c_int]
    L.llsm_b200_halo_length.argtypes = [C.POINTER(abi.Conf)]
    L.llsm_b200_frame_position.argtypes = [C.c_int, C.c_float, C.c_float]
    L.llsm_b200_synthesize_harmonics.argtypes = [P, C.POINTER(abi.Conf), C.POINTER(abi.Frames),
                                                 C.POINTER(abi.SOptions), P, C.c_int, C.c_int]
    L.llsm_b200_analyze_l0.argtypes = [P, C.POINTER(abi.Conf), C.POINTER(abi.AOptions), P, C.c_int,
                                       C.c_int, C.POINTER(abi.FramesOut), P]
    L.llsm_b200_analyze_l0_host.argtypes = L.llsm_b200_analyze_l0.argtypes
    L.llsm_b200_anasynth_host.argtypes = [P, C.POINTER(abi.Conf), C.POINTER(abi.AOptions), C.POINTER(abi.SOptions), P,
                                          C.c_int, C.c_int, P, P, C.c_int, C.POINTER(abi.Output)]
    L.llsm_b200_tolayer1.argtypes = [P, C.POINTER(abi.Conf), C.POINTER(abi.Frames), C.c_int, C.POINTER(abi.Layer1)]
    L.llsm_b200_synthesize_l1.argtypes = [P, C.POINTER(abi.Conf), C.POINTER(abi.Frames), C.POINTER(abi.Layer1), P,
                                          C.POINTER(abi.SOptions), C.POINTER(abi.Output)]
    L.llsm_b200_tolayer0.argtypes = [P, C.POINTER(abi.Conf), P, P, C.POINTER(abi.Layer1), P, P, P]
    L.llsm_b200_rt_template_length.argtypes = [C.c_float]
    L.llsm_b200_rt_fft_size.argtypes = [C.c_float, C.c_float]
    L.llsm_b200_rt_create.argtypes = [P, C.POINTER(abi.Conf), C.POINTER(abi.SOptions), C.c_int, C.POINTER(C.c_void_p)]
    L.llsm_b200_rt_destroy.argtypes = [P]
    L.llsm_b200_rt_latency.argtypes = [P]
    L.llsm_b200_rt_output_length.argtypes = [P, C.c_int]
    L.llsm_b200_rt_clear.argtypes = [P]
    L.llsm_b200_rt_feed.argtypes = [P, C.POINTER(abi.Frames), C.c_int, P, P, C.c_int, C.POINTER(C.c_int)]
    L.llsm_b200_rt_feed_host.argtypes = L.llsm_b200_rt_feed.argtypes
    L.llsm_b200_rt_create_l1.argtypes = [P, C.POINTER(abi.Conf), C.POINTER(abi.SOptions), C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_void_p)]
    L.llsm_b200_rt_feed_l1.argtypes = [P, C.POINTER(abi.Frames), C.POINTER(abi.Layer1), P, C.c_int, P, P, C.c_int,
                                       C.POINTER(C.c_int)]
    L.llsm_b200_rt_feed_l1_host.argtypes = [P, C.POINTER(abi.Frames), C.POINTER(abi.Layer1), P, C.c_int, P, P, P, P,
                                            C.c_int, C.POINTER(C.c_int)]
    L.llsm_b200_chunk_phasepropagate.argtypes = [P, C.POINTER(abi.Conf), P, C.POINTER(abi.FramesOut),
                                                 C.POINTER(abi.Layer1), C.c_int]
    L.llsm_b200_chunk_phasesync_rps.argtypes = L.llsm_b200_chunk_phasepropagate.argtypes
    L.llsm_b200_coder_dimension.argtypes = [C.c_int, C.c_int]
    L.llsm_b200_coder_encode.argtypes = [P, C.POINTER(abi.Conf), P, P, P, C.POINTER(abi.Layer1), C.c_int, C.c_int, P]
    L.llsm_b200_coder_decode.argtypes = [P, C.POINTER(abi.Conf), P, P, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(abi.FramesOut), C.POINTER(abi.Layer1)]
    L.llsm_b200_frames_stretch.argtypes = [P, C.POINTER(abi.Conf), C.POINTER(abi.Frames), C.POINTER(abi.Layer1), C.c_int,
                                           P, P, P, C.c_int, C.POINTER(abi.FramesOut), C.POINTER(abi.Layer1)]
    L.llsm_b200_stretch_map.argtypes = [C.c_int, C.c_int, P, P, P]
    L.llsm_b200_frames_blob_size.restype = C.c_size_t
    L.llsm_b200_frames_blob_size.argtypes = [C.POINTER(abi.Conf), C.POINTER(abi.Frames)]
    L.llsm_b200_frames_pack.argtypes = [C.POINTER(abi.Conf), C.POINTER(abi.Frames), P, C.c_size_t]
    L.llsm_b200_frames_unpack.argtypes = [P, C.c_size_t, C.POINTER(abi.Conf), C.POINTER(abi.Frames)]
    L.llsm_b200_comm_unique_id.argtypes = [P]
    L.llsm_b200_comm_init.argtypes = [P, C.c_char_p, C.c_int, C.c_int]
    L.llsm_b200_comm_attach.argtypes = [P, P, C.c_int, C.c_int]
    L.llsm_b200_comm_destroy.argtypes = [P]
    L.llsm_b200_shard_position.argtypes = [C.POINTER(abi.Conf), C.c_int]
    L.llsm_b200_halo_exchange.argtypes = [P, C.POINTER(abi.Conf), C.POINTER(abi.Output), P]
    L.llsm_b200_set_kernel_timing.argtypes = [P, C.c_int]
    L.llsm_b200_kernel_timing_read.argtypes = [P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise LlsmB200Error("libllsm2_b200 error %d: %s" % (rc, lib().llsm_b200_last_error().decode()))
