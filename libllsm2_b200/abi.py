"""ctypes mirror of include/llsm_b200.h (the C ABI of the hot path)."""
import ctypes as C

MAXCHANNEL = 8
c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)


class Conf(C.Structure):
    _fields_ = [("nutt", C.c_int), ("nfrm", C.c_int), ("maxnhar", C.c_int), ("maxnhar_e", C.c_int),
                ("npsd", C.c_int), ("nchannel", C.c_int), ("fs", C.c_float), ("thop", C.c_float),
                ("chanfreq", C.c_float * MAXCHANNEL), ("lip_radius", C.c_float)]


class Frames(C.Structure):
    _fields_ = [("nfrm_utt", C.c_void_p), ("f0", C.c_void_p), ("nhar", C.c_void_p),
                ("ampl", C.c_void_p), ("phse", C.c_void_p), ("psd", C.c_void_p),
                ("psdres", C.c_void_p), ("edc", C.c_void_p), ("enhar", C.c_void_p),
                ("eampl", C.c_void_p), ("ephse", C.c_void_p)]


class FramesOut(C.Structure):
    _fields_ = [("f0", C.c_void_p), ("nhar", C.c_void_p), ("ampl", C.c_void_p), ("phse", C.c_void_p),
                ("psd", C.c_void_p), ("psdres", C.c_void_p), ("edc", C.c_void_p),
                ("enhar", C.c_void_p), ("eampl", C.c_void_p), ("ephse", C.c_void_p)]


class SOptions(C.Structure):
    _fields_ = [("use_iczt", C.c_int), ("iczt_param_a", C.c_float), ("iczt_param_b", C.c_float),
                ("white", C.c_void_p), ("seed", C.c_uint64)]


class Output(C.Structure):
    _fields_ = [("y", C.c_void_p), ("y_sin", C.c_void_p), ("y_noise", C.c_void_p),
                ("stride", C.c_int)]


class AOptions(C.Structure):
    _fields_ = [("f0_refine", C.c_int), ("hm_method", C.c_int), ("rel_winsize", C.c_float)]


def make_conf(nutt, nfrm, maxnhar, maxnhar_e, npsd, nchannel, fs, thop,
              chanfreq=(2000.0, 4000.0, 8000.0), lip_radius=1.5):
    c = Conf()
    c.nutt, c.nfrm, c.maxnhar, c.maxnhar_e = nutt, nfrm, maxnhar, maxnhar_e
    c.npsd, c.nchannel, c.fs, c.thop, c.lip_radius = npsd, nchannel, fs, thop, lip_radius
    for i, f in enumerate(chanfreq[:max(nchannel - 1, 0)]):
        c.chanfreq[i] = f
    return c


def default_soptions(white_ptr=None, seed=0):
    """llsm_create_soptions defaults (reference layer0.c:78-87)."""
    o = SOptions()
    o.use_iczt, o.iczt_param_a, o.iczt_param_b = 1, 0.275, 2.26
    o.white, o.seed = white_ptr, seed
    return o


class Layer1(C.Structure):
    _fields_ = [("rd", C.c_void_p), ("vtmagn", C.c_void_p), ("vsphse", C.c_void_p), ("nvs", C.c_void_p),
                ("nspec", C.c_int)]
