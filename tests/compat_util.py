"""ctypes view of the llsm.h data model, usable with BOTH the reference build (oracle/_ref) and the
drop-in library (libllsm2_b200.so): the two export the same C API."""
import ctypes as C
import numpy as np

fp = C.POINTER(C.c_float)


class HM(C.Structure):
    _fields_ = [("ampl", fp), ("phse", fp), ("nhar", C.c_int)]


class NM(C.Structure):
    _fields_ = [("eenv", C.POINTER(C.POINTER(HM))), ("edc", fp), ("psd", fp), ("npsd", C.c_int), ("nchannel", C.c_int)]


class Container(C.Structure):
    _fields_ = [("members", C.POINTER(C.c_void_p)), ("destructors", C.POINTER(C.c_void_p)),
                ("copyctors", C.POINTER(C.c_void_p)), ("nmember", C.c_int)]


class Chunk(C.Structure):
    _fields_ = [("conf", C.POINTER(Container)), ("frames", C.POINTER(C.POINTER(Container)))]


class Output(C.Structure):
    _fields_ = [("ny", C.c_int), ("fs", C.c_float), ("y", fp), ("y_sin", fp), ("y_noise", fp)]


class AOptions(C.Structure):
    _fields_ = [("thop", C.c_float), ("maxnhar", C.c_int), ("maxnhar_e", C.c_int), ("npsd", C.c_int),
                ("nchannel", C.c_int), ("chanfreq", fp), ("lip_radius", C.c_float), ("f0_refine", C.c_int),
                ("hm_method", C.c_int), ("rel_winsize", C.c_float)]


class SOptions(C.Structure):
    _fields_ = [("fs", C.c_float), ("use_iczt", C.c_int), ("use_l1", C.c_int), ("iczt_param_a", C.c_float),
                ("iczt_param_b", C.c_float)]


F0, HMI, NMI, PSDRES = 0, 1, 2, 3
CONF_NFRM = 0


def bind(L):
    P = C.c_void_p
    L.llsm_create_aoptions.restype = C.POINTER(AOptions)
    L.llsm_aoptions_toconf.restype = C.POINTER(Container)
    L.llsm_aoptions_toconf.argtypes = [C.POINTER(AOptions), C.c_float]
    L.llsm_create_chunk.restype = C.POINTER(Chunk)
    L.llsm_create_chunk.argtypes = [C.POINTER(Container), C.c_int]
    L.llsm_copy_chunk.restype = C.POINTER(Chunk)
    L.llsm_container_get.restype = P
    L.llsm_container_get.argtypes = [C.POINTER(Container), C.c_int]
    L.llsm_container_attach_.argtypes = [C.POINTER(Container), C.c_int, P, P, P]
    L.llsm_create_hmframe.restype = C.POINTER(HM)
    L.llsm_copy_hmframe.restype = C.POINTER(HM)
    L.llsm_create_nmframe.restype = C.POINTER(NM)
    L.llsm_copy_nmframe.restype = C.POINTER(NM)
    L.llsm_create_fparray.restype = fp
    L.llsm_copy_fparray.restype = fp
    L.llsm_create_container.restype = C.POINTER(Container)
    L.llsm_copy_container.restype = C.POINTER(Container)
    L.llsm_create_soptions.restype = C.POINTER(SOptions)
    L.llsm_create_soptions.argtypes = [C.c_float]
    L.llsm_synthesize.restype = C.POINTER(Output)
    L.llsm_analyze.restype = C.POINTER(Chunk)
    L.llsm_analyze.argtypes = [C.POINTER(AOptions), fp, C.c_int, C.c_float, fp, C.c_int, C.POINTER(fp)]
    L.llsm_chunk_getf0.restype = fp
    L.llsm_frame_phaseshift.argtypes = [C.POINTER(Container), C.c_float]
    L.llsm_hmframe_phaseshift.argtypes = [C.POINTER(HM), C.c_float]
    L.llsm_create_frame.restype = C.POINTER(Container)
    L.llsm_create_fp.restype = fp
    L.llsm_create_fp.argtypes = [C.c_float]
    L.llsm_create_int.restype = C.POINTER(C.c_int)
    return L


def fn_ptr(L, name):
    return C.cast(getattr(L, name), C.c_void_p)


def build_chunk(L, frames, conf, b=0):
    """Build an llsm_chunk for utterance b from flat arrays through the public API of library L."""
    ao = L.llsm_create_aoptions()
    ao.contents.thop = conf.thop; ao.contents.maxnhar = conf.maxnhar; ao.contents.maxnhar_e = conf.maxnhar_e
    ao.contents.npsd = conf.npsd; ao.contents.nchannel = conf.nchannel
    for i in range(conf.nchannel - 1):
        ao.contents.chanfreq[i] = conf.chanfreq[i]
    cf = L.llsm_aoptions_toconf(ao, C.c_float(conf.fs / 2.0))
    C.cast(L.llsm_container_get(cf, CONF_NFRM), C.POINTER(C.c_int))[0] = conf.nfrm
    ck = L.llsm_create_chunk(cf, 1)
    L.llsm_delete_container(cf)
    L.llsm_delete_aoptions(ao)
    for i in range(conf.nfrm):
        fr = ck.contents.frames[i]
        C.cast(L.llsm_container_get(fr, F0), fp)[0] = float(frames["f0"][b, i])
        nh = int(frames["nhar"][b, i])
        if frames["f0"][b, i] > 0:
            hm = L.llsm_create_hmframe(nh)
            for k in range(nh):
                hm.contents.ampl[k] = float(frames["ampl"][b, i, k]); hm.contents.phse[k] = float(frames["phse"][b, i, k])
            L.llsm_container_attach_(fr, HMI, C.cast(hm, C.c_void_p), fn_ptr(L, "llsm_delete_hmframe"), fn_ptr(L, "llsm_copy_hmframe"))
        nm = C.cast(L.llsm_container_get(fr, NMI), C.POINTER(NM))
        for j in range(conf.npsd):
            nm.contents.psd[j] = float(frames["psd"][b, i, j])
        for c in range(conf.nchannel):
            nm.contents.edc[c] = float(frames["edc"][b, i, c])
            ne = int(frames["enhar"][b, i, c])
            e = L.llsm_create_hmframe(ne)
            for k in range(ne):
                e.contents.ampl[k] = float(frames["eampl"][b, i, c, k]); e.contents.phse[k] = float(frames["ephse"][b, i, c, k])
            L.llsm_copy_hmframe_inplace(nm.contents.eenv[c], e)
            L.llsm_delete_hmframe(e)
        if frames.get("psdres") is not None:
            r = L.llsm_create_fparray(conf.npsd)
            for j in range(conf.npsd):
                r[j] = float(frames["psdres"][b, i, j])
            L.llsm_container_attach_(fr, PSDRES, C.cast(r, C.c_void_p), fn_ptr(L, "llsm_delete_fparray"), fn_ptr(L, "llsm_copy_fparray"))
    return ck


def output_arrays(o):
    ny = o.contents.ny
    g = lambda p: np.ctypeslib.as_array(p, (ny,)).copy()
    return g(o.contents.y), g(o.contents.y_sin), g(o.contents.y_noise)


def chunk_to_flat(L, ck, conf):
    """Read a chunk produced by llsm_analyze back into flat arrays."""
    F = conf.nfrm
    o = dict(f0=np.zeros(F, np.float32), nhar=np.zeros(F, np.int32), ampl=np.zeros((F, conf.maxnhar), np.float32),
             phse=np.zeros((F, conf.maxnhar), np.float32), psd=np.zeros((F, conf.npsd), np.float32),
             psdres=np.zeros((F, conf.npsd), np.float32), edc=np.zeros((F, conf.nchannel), np.float32),
             enhar=np.zeros((F, conf.nchannel), np.int32))
    for i in range(F):
        fr = ck.contents.frames[i]
        o["f0"][i] = C.cast(L.llsm_container_get(fr, F0), fp)[0]
        hm = L.llsm_container_get(fr, HMI)
        if hm and o["f0"][i] > 0:
            hm = C.cast(hm, C.POINTER(HM)).contents
            o["nhar"][i] = hm.nhar
            o["ampl"][i, :hm.nhar] = np.ctypeslib.as_array(hm.ampl, (hm.nhar,)) if hm.nhar else 0
            o["phse"][i, :hm.nhar] = np.ctypeslib.as_array(hm.phse, (hm.nhar,)) if hm.nhar else 0
        nm = C.cast(L.llsm_container_get(fr, NMI), C.POINTER(NM)).contents
        o["psd"][i] = np.ctypeslib.as_array(nm.psd, (conf.npsd,))
        o["edc"][i] = np.ctypeslib.as_array(nm.edc, (conf.nchannel,))
        for c in range(conf.nchannel):
            o["enhar"][i, c] = nm.eenv[c].contents.nhar
        r = L.llsm_container_get(fr, PSDRES)
        if r:
            o["psdres"][i] = np.ctypeslib.as_array(C.cast(r, fp), (conf.npsd,))
    return o
