"""CPU test (no GPU) of the flow.cu kernels' per-thread arithmetic: tests/emul/flow_emul.cpp compiles
csrc/flow_device.h + csrc/flow_setup.h (the code the kernels are thin wrappers of) as host C++ and runs it thread
by thread; the result must equal the reference's goldens bit for bit.  This is test infrastructure: the product has
no CPU path (the GPU parity tests in test_gpu_flow.py are the real check of the kernels)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import mvtools_b200 as mb
import pipeline_flow

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_flow.json")))
_LIB = None


def emul():
    global _LIB
    if _LIB is None:
        d = os.path.join(HERE, "emul")
        so = os.path.join(d, "libflow_emul.so")
        pkg = os.path.join(HERE, "..", "vapoursynth-mvtools_b200")
        deps = [os.path.join(d, "flow_emul.cpp"), os.path.join(pkg, "csrc", "flow_device.h"), os.path.join(pkg, "csrc", "flow_setup.h")]
        if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-fvisibility=hidden", "-Wall", "-ffp-contract=off", "-shared", "-o", so, deps[0],
                                   "-L" + pkg, "-lmvtools_b200", "-Wl,-rpath," + os.path.abspath(pkg)])
        mb.lib()  # libmvtools_b200.so (host logic only) must be loaded first
        L = C.CDLL(so)
        vp, PP = C.c_void_p, C.POINTER
        L.emul_scdetection_frame.argtypes = [PP(mb.ScdParams), vp, PP(C.c_int32)]
        L.emul_mask_frame.argtypes = [PP(mb.MaskParams), vp, C.c_ssize_t, vp, PP(vp), PP(C.c_ssize_t)]
        L.emul_finest_frame.argtypes = [PP(mb.SuperParams), PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t)]
        L.emul_flow_frame.argtypes = [PP(mb.FlowParams), PP(mb.SuperParams), PP(vp), PP(C.c_ssize_t), vp, C.c_int, PP(vp), PP(C.c_ssize_t), PP(C.c_int32)]
        L.emul_flowinter_frame.argtypes = [PP(mb.FlowInterParams), PP(mb.SuperParams), C.c_int, PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t), vp, vp, vp, vp, PP(vp), PP(C.c_ssize_t), PP(C.c_int32)]
        L.emul_blend_frame.argtypes = [PP(mb.SuperParams), C.c_int, PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t)]
        L.emul_flowblur_frame.argtypes = [PP(mb.FlowBlurParams), PP(mb.SuperParams), PP(vp), PP(C.c_ssize_t), vp, vp, PP(vp), PP(C.c_ssize_t), PP(C.c_int32)]
        L.emul_blockfps_frame.argtypes = [PP(mb.BlockFPSParams), PP(mb.SuperParams), C.c_int, PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t), vp, vp, PP(vp), PP(C.c_ssize_t), PP(C.c_int32)]
        _LIB = L
    return _LIB


def scd(scp, vectors):
    out = C.c_int32(0)
    assert emul().emul_scdetection_frame(C.byref(scp), vectors.ctypes.data, C.byref(out)) == 0
    return out.value


def mask(mp, src_luma, vectors):
    dst = mb.alloc_planes([mp.nWidth, mp.nWidthUV, mp.nWidthUV], [mp.nHeight, mp.nHeightUV, mp.nHeightUV], 8, fill=0xA5)
    d, ds = mb.plane_ptrs(dst)
    assert emul().emul_mask_frame(C.byref(mp), src_luma.ctypes.data, src_luma.strides[0], vectors.ctypes.data, d, ds) == 0
    return dst


def finest(sp, super_planes):
    l = mb.finest_layout(sp)
    dst = mb.alloc_planes(list(l.plane_w)[: sp.nplanes], list(l.plane_h)[: sp.nplanes], sp.bits, fill=0xA5)
    s, ss = mb.plane_ptrs(super_planes)
    d, ds = mb.plane_ptrs(dst)
    assert emul().emul_finest_frame(C.byref(sp), s, ss, d, ds) == 0
    return dst


def flow(fp, sp, ref_super, vectors, field_shift=0):
    w = [fp.vd.nWidth, fp.nWidthUV, fp.nWidthUV][: fp.nplanes]
    h = [fp.vd.nHeight, fp.nHeightUV, fp.nHeightUV][: fp.nplanes]
    dst = mb.alloc_planes(w, h, fp.bits, fill=0xA5)
    r, rs = (None, None) if ref_super is None else mb.plane_ptrs(ref_super)
    d, ds = mb.plane_ptrs(dst)
    usable = C.c_int32(0)
    assert emul().emul_flow_frame(C.byref(fp), C.byref(sp), r, rs, vectors.ctypes.data, field_shift, d, ds, C.byref(usable)) == 0
    return dst if usable.value else None


def inter(p, sp, time256, left_super, right_super, vec_b, vec_f, vec_bb=None, vec_ff=None):
    w = [p.vd.nWidth, p.nWidthUV, p.nWidthUV][: p.nplanes]
    h = [p.vd.nHeight, p.nHeightUV, p.nHeightUV][: p.nplanes]
    dst = mb.alloc_planes(w, h, p.bits, fill=0xA5)
    l, ls = mb.plane_ptrs(left_super)
    r, rs = mb.plane_ptrs(right_super)
    d, ds = mb.plane_ptrs(dst)
    usable = C.c_int32(0)
    ptr = lambda v: None if v is None else v.ctypes.data
    assert emul().emul_flowinter_frame(C.byref(p), C.byref(sp), time256, l, ls, r, rs, ptr(vec_b), ptr(vec_f), ptr(vec_bb), ptr(vec_ff), d, ds, C.byref(usable)) == 0
    return dst if usable.value else None


def blur(p, sp, super_planes, vec_b, vec_f):
    w = [p.vd.nWidth, p.nWidthUV, p.nWidthUV][: p.nplanes]
    h = [p.vd.nHeight, p.nHeightUV, p.nHeightUV][: p.nplanes]
    dst = mb.alloc_planes(w, h, p.bits, fill=0xA5)
    s, ss = mb.plane_ptrs(super_planes)
    d, ds = mb.plane_ptrs(dst)
    usable = C.c_int32(0)
    ptr = lambda v: None if v is None else v.ctypes.data
    assert emul().emul_flowblur_frame(C.byref(p), C.byref(sp), s, ss, ptr(vec_b), ptr(vec_f), d, ds, C.byref(usable)) == 0
    return dst if usable.value else None


def bfps(p, sp, time256, left_super, right_super, vec_b, vec_f):
    w = [p.vd.nWidth, p.nWidthUV, p.nWidthUV][: p.nplanes]
    h = [p.vd.nHeight, p.nHeightUV, p.nHeightUV][: p.nplanes]
    dst = mb.alloc_planes(w, h, p.bits, fill=0xA5)
    l, ls = mb.plane_ptrs(left_super)
    r, rs = mb.plane_ptrs(right_super)
    d, ds = mb.plane_ptrs(dst)
    usable = C.c_int32(0)
    assert emul().emul_blockfps_frame(C.byref(p), C.byref(sp), time256, l, ls, r, rs, vec_b.ctypes.data, vec_f.ctypes.data, d, ds, C.byref(usable)) == 0
    return dst if usable.value else None


def blend(sp, time256, src, ref):
    dst = mb.alloc_planes([a.shape[1] for a in src], [a.shape[0] for a in src], sp.bits, fill=0xA5)
    s, ss = mb.plane_ptrs(src)
    r, rs = mb.plane_ptrs(ref)
    d, ds = mb.plane_ptrs(dst)
    assert emul().emul_blend_frame(C.byref(sp), time256, s, ss, r, rs, d, ds) == 0
    return dst


@pytest.mark.parametrize("case", sorted(pipeline_flow.FLOW))
def test_kernel_arithmetic_matches_reference_goldens(built, case):
    got = pipeline_flow.run_engine_flow(case, scd, mask, finest, flow)
    got["inter"] = pipeline_flow.run_engine_inter(case, inter, blend)
    got["fps"] = pipeline_flow.run_engine_fps(case, inter, blend)
    got["blur"] = pipeline_flow.run_engine_blur(case, blur)
    got["bfps"] = pipeline_flow.run_engine_bfps(case, bfps, blend)
    for stage in ("scd", "mask", "flow", "finest", "inter", "fps", "blur", "bfps"):
        assert got[stage] == GOLD[case][stage], "%s: emulated kernels' %s differs from the reference's golden hashes" % (case, stage)


def test_field_shift_matches_oracle(built):
    """mv.Flow's fields=True path (a vertical shift added to every vector, MVFlow.cpp:301-305) against the oracle"""
    import oracle
    case = "blk8_ov4_pel2_p8"
    frames, sups, vecs = pipeline_flow.inputs(case)
    sp, aps, P = pipeline_flow.flow_params_for(case)
    for (v, fp) in P["flow"][:3]:
        for fs in (1, -1):
            a = flow(fp, sp, sups[2], vecs[v][1], fs)
            b = oracle.flow_frame(fp, sp, sups[2], vecs[v][1], fs)
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
