"""TEST INFRASTRUCTURE: ctypes binding of the CPU oracle (oracle/_ref/liboracle.so).
Takes the same POD parameter structs as the C ABI (vapoursynth-mvtools_b200/mvtools_b200.py)."""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "..", "vapoursynth-mvtools_b200"))
import mvtools_b200 as mb  # noqa: E402  (struct definitions only)

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(os.path.join(_HERE, "_ref", "liboracle.so"))
        L.mvo_last_error.restype = C.c_char_p
        vp, PP = C.c_void_p, C.POINTER
        L.mvo_super_frame.argtypes = [PP(mb.SuperParams), PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t)]
        L.mvo_super_frame_pel.argtypes = [PP(mb.SuperParams), PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t), C.c_int, PP(vp), PP(C.c_ssize_t)]
        L.mvo_analyse_frame.argtypes = [PP(mb.AnalyseParams), PP(mb.SuperParams), PP(vp), PP(C.c_ssize_t), PP(vp),
                                        PP(C.c_ssize_t), C.c_int, vp]
        L.mvo_degrain_frame.argtypes = [PP(mb.DegrainParams), PP(mb.SuperParams), PP(vp), PP(C.c_ssize_t), PP(vp),
                                        PP(C.c_ssize_t), PP(vp), PP(vp), PP(C.c_ssize_t)]
        L.mvo_recalculate_frame.argtypes = [PP(mb.RecalculateParams), PP(mb.SuperParams), PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t), vp, C.c_int, vp]
        L.mvo_compensate_frame.argtypes = [PP(mb.CompensateParams), PP(mb.SuperParams), PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t), vp, C.c_int,
                                           PP(vp), PP(C.c_ssize_t)]
        L.mvo_overlap_windows.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.mvo_overlap_windows.restype = None
        L.mvo_scdetection_frame.argtypes = [PP(mb.ScdParams), vp, PP(C.c_int32)]
        L.mvo_mask_frame.argtypes = [PP(mb.MaskParams), vp, C.c_ssize_t, vp, PP(vp), PP(C.c_ssize_t)]
        L.mvo_finest_frame.argtypes = [PP(mb.SuperParams), PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t)]
        L.mvo_flow_frame.argtypes = [PP(mb.FlowParams), PP(mb.SuperParams), PP(vp), PP(C.c_ssize_t), vp, C.c_int, PP(vp), PP(C.c_ssize_t), PP(C.c_int32)]
        L.mvo_flowinter_frame.argtypes = [PP(mb.FlowInterParams), PP(mb.SuperParams), C.c_int, PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t), vp, vp, vp, vp, PP(vp), PP(C.c_ssize_t), PP(C.c_int32)]
        L.mvo_blend_frame.argtypes = [PP(mb.SuperParams), C.c_int, PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t)]
        L.mvo_flowblur_frame.argtypes = [PP(mb.FlowBlurParams), PP(mb.SuperParams), PP(vp), PP(C.c_ssize_t), vp, vp, PP(vp), PP(C.c_ssize_t), PP(C.c_int32)]
        L.mvo_blockfps_frame.argtypes = [PP(mb.BlockFPSParams), PP(mb.SuperParams), C.c_int, PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t), vp, vp, PP(vp), PP(C.c_ssize_t), PP(C.c_int32)]
        L.mvo_simple_resize_tables.argtypes = [C.c_int, C.c_int, vp, vp]
        L.mvo_simple_resize_tables.restype = None
        for n in ("mvo_sad", "mvo_satd"):
            f = getattr(L, n)
            f.restype = C.c_uint
            f.argtypes = [vp, C.c_ssize_t, vp, C.c_ssize_t, C.c_int, C.c_int, C.c_int]
        _LIB = L
    return _LIB


def _check(rc):
    if rc:
        raise RuntimeError(lib().mvo_last_error().decode())


def super_frame(sp, src_planes, pel_planes=None, pel_padded=False):
    if pel_planes is not None:
        return super_frame_pel(sp, src_planes, pel_planes, pel_padded)
    g = mb.SuperGeom()
    # geometry comes with the params; plane sizes follow MVSuper.c:254-262
    pw = [sp.super_width] + [sp.super_width >> sp.ssw] * 2
    ph = [sp.super_height] + [sp.super_height >> sp.ssh] * 2
    dst = mb.alloc_planes(pw[: sp.nplanes], ph[: sp.nplanes], sp.bits, fill=0xA5)
    s, ss = mb.plane_ptrs(src_planes)
    d, ds = mb.plane_ptrs(dst)
    _check(lib().mvo_super_frame(C.byref(sp), s, ss, d, ds))
    return dst


def super_frame_pel(sp, src_planes, pel_planes, pel_padded):
    pw = [sp.super_width] + [sp.super_width >> sp.ssw] * 2
    ph = [sp.super_height] + [sp.super_height >> sp.ssh] * 2
    dst = mb.alloc_planes(pw[: sp.nplanes], ph[: sp.nplanes], sp.bits, fill=0xA5)
    s, ss = mb.plane_ptrs(src_planes)
    q, qs = mb.plane_ptrs(pel_planes)
    d, ds = mb.plane_ptrs(dst)
    _check(lib().mvo_super_frame_pel(C.byref(sp), s, ss, q, qs, int(pel_padded), d, ds))
    return dst


def analyse_frame(ap, sp, src_super, ref_super, field_shift=0):
    out = np.zeros(ap.vectors_size, dtype=np.uint8)
    s, ss = mb.plane_ptrs(src_super)
    if ref_super is None:
        r, rs = None, None
    else:
        r, rs = mb.plane_ptrs(ref_super)
    _check(lib().mvo_analyse_frame(C.byref(ap), C.byref(sp), s, ss, r, rs, field_shift, out.ctypes.data))
    return out


def degrain_frame(dp, sp, src_planes, ref_supers, vectors):
    n = 2 * dp.radius
    refs = (C.c_void_p * (3 * n))()
    rstr = (C.c_ssize_t * (3 * n))()
    for r in range(n):
        if ref_supers[r] is None:
            continue
        for p, a in enumerate(ref_supers[r]):
            refs[r * 3 + p] = a.ctypes.data
            rstr[r * 3 + p] = a.strides[0]
    vecs = (C.c_void_p * n)(*[v.ctypes.data for v in vectors])
    dst = mb.alloc_planes([a.shape[1] for a in src_planes], [a.shape[0] for a in src_planes], dp.bits, fill=0xA5)
    s, ss = mb.plane_ptrs(src_planes)
    d, ds = mb.plane_ptrs(dst)
    _check(lib().mvo_degrain_frame(C.byref(dp), C.byref(sp), s, ss, refs, rstr, vecs, d, ds))
    return dst


def recalculate_frame(rp, sp, src_super, ref_super, old_vectors, field_shift=0):
    out = np.zeros(rp.vectors_size, dtype=np.uint8)
    s, ss = mb.plane_ptrs(src_super)
    r, rs = (None, None) if ref_super is None else mb.plane_ptrs(ref_super)
    _check(lib().mvo_recalculate_frame(C.byref(rp), C.byref(sp), s, ss, r, rs, old_vectors.ctypes.data, field_shift, out.ctypes.data))
    return out


def compensate_frame(cp, sp, src_super, ref_super, vectors, field_shift=0):
    dst = mb.alloc_planes(list(cp.nWidth)[: cp.nplanes], list(cp.nHeight)[: cp.nplanes], cp.bits, fill=0xA5)
    s, ss = mb.plane_ptrs(src_super)
    r, rs = (None, None) if ref_super is None else mb.plane_ptrs(ref_super)
    d, ds = mb.plane_ptrs(dst)
    _check(lib().mvo_compensate_frame(C.byref(cp), C.byref(sp), s, ss, r, rs, vectors.ctypes.data, field_shift, d, ds))
    return dst


def overlap_windows(nx, ny, ox, oy):
    out = np.zeros((9, ny, nx), dtype=np.int16)
    lib().mvo_overlap_windows(nx, ny, ox, oy, out.ctypes.data)
    return out


# ---- consumers of the vector clips (mvtools_oracle_flow.c)

def scdetection_frame(scp, vectors):
    out = C.c_int32(0)
    _check(lib().mvo_scdetection_frame(C.byref(scp), vectors.ctypes.data, C.byref(out)))
    return out.value


def mask_frame(mp, src_luma, vectors):
    dst = mb.alloc_planes([mp.nWidth, mp.nWidthUV, mp.nWidthUV], [mp.nHeight, mp.nHeightUV, mp.nHeightUV], 8, fill=0xA5)
    d, ds = mb.plane_ptrs(dst)
    sl = None if src_luma is None else src_luma.ctypes.data
    _check(lib().mvo_mask_frame(C.byref(mp), sl, 0 if src_luma is None else src_luma.strides[0], vectors.ctypes.data, d, ds))
    return dst


def finest_frame(sp, super_planes):
    l = mb.finest_layout(sp)
    dst = mb.alloc_planes(list(l.plane_w)[: sp.nplanes], list(l.plane_h)[: sp.nplanes], sp.bits, fill=0xA5)
    s, ss = mb.plane_ptrs(super_planes)
    d, ds = mb.plane_ptrs(dst)
    _check(lib().mvo_finest_frame(C.byref(sp), s, ss, d, ds))
    return dst


def flow_frame(fp, sp, ref_super, vectors, field_shift=0):
    w = [fp.vd.nWidth, fp.nWidthUV, fp.nWidthUV][: fp.nplanes]
    h = [fp.vd.nHeight, fp.nHeightUV, fp.nHeightUV][: fp.nplanes]
    dst = mb.alloc_planes(w, h, fp.bits, fill=0xA5)
    r, rs = (None, None) if ref_super is None else mb.plane_ptrs(ref_super)
    d, ds = mb.plane_ptrs(dst)
    usable = C.c_int32(0)
    _check(lib().mvo_flow_frame(C.byref(fp), C.byref(sp), r, rs, vectors.ctypes.data, field_shift, d, ds, C.byref(usable)))
    return dst if usable.value else None


def simple_resize_tables(out, in_):
    off = np.zeros(out, dtype=np.int32)
    wt = np.zeros(out, dtype=np.int32)
    lib().mvo_simple_resize_tables(out, in_, off.ctypes.data, wt.ctypes.data)
    return off, wt


def flowinter_frame(p, sp, time256, left_super, right_super, vec_b, vec_f, vec_bb=None, vec_ff=None):
    w = [p.vd.nWidth, p.nWidthUV, p.nWidthUV][: p.nplanes]
    h = [p.vd.nHeight, p.nHeightUV, p.nHeightUV][: p.nplanes]
    dst = mb.alloc_planes(w, h, p.bits, fill=0xA5)
    l, ls = mb.plane_ptrs(left_super)
    r, rs = mb.plane_ptrs(right_super)
    d, ds = mb.plane_ptrs(dst)
    usable = C.c_int32(0)
    ptr = lambda v: None if v is None else v.ctypes.data
    _check(lib().mvo_flowinter_frame(C.byref(p), C.byref(sp), time256, l, ls, r, rs, ptr(vec_b), ptr(vec_f), ptr(vec_bb), ptr(vec_ff), d, ds, C.byref(usable)))
    return dst if usable.value else None


def blend_frame(sp, time256, src, ref):
    dst = mb.alloc_planes([a.shape[1] for a in src], [a.shape[0] for a in src], sp.bits, fill=0xA5)
    s, ss = mb.plane_ptrs(src)
    r, rs = mb.plane_ptrs(ref)
    d, ds = mb.plane_ptrs(dst)
    _check(lib().mvo_blend_frame(C.byref(sp), time256, s, ss, r, rs, d, ds))
    return dst


def flowblur_frame(p, sp, super_planes, vec_b, vec_f):
    w = [p.vd.nWidth, p.nWidthUV, p.nWidthUV][: p.nplanes]
    h = [p.vd.nHeight, p.nHeightUV, p.nHeightUV][: p.nplanes]
    dst = mb.alloc_planes(w, h, p.bits, fill=0xA5)
    s, ss = mb.plane_ptrs(super_planes)
    d, ds = mb.plane_ptrs(dst)
    usable = C.c_int32(0)
    ptr = lambda v: None if v is None else v.ctypes.data
    _check(lib().mvo_flowblur_frame(C.byref(p), C.byref(sp), s, ss, ptr(vec_b), ptr(vec_f), d, ds, C.byref(usable)))
    return dst if usable.value else None


def blockfps_frame(p, sp, time256, left_super, right_super, vec_b, vec_f):
    w = [p.vd.nWidth, p.nWidthUV, p.nWidthUV][: p.nplanes]
    h = [p.vd.nHeight, p.nHeightUV, p.nHeightUV][: p.nplanes]
    dst = mb.alloc_planes(w, h, p.bits, fill=0xA5)
    l, ls = mb.plane_ptrs(left_super)
    r, rs = mb.plane_ptrs(right_super)
    d, ds = mb.plane_ptrs(dst)
    usable = C.c_int32(0)
    _check(lib().mvo_blockfps_frame(C.byref(p), C.byref(sp), time256, l, ls, r, rs, vec_b.ctypes.data, vec_f.ctypes.data, d, ds, C.byref(usable)))
    return dst if usable.value else None
