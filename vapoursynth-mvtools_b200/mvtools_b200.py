"""ctypes binding of include/mvtools_b200.h (libmvtools_b200.so).

This is the thin Python face of the C ABI: the structs below mirror the header field
for field, and `lib()` loads the in-tree shared library.  There is no Python or CPU
implementation behind it -- if the CUDA library cannot be loaded, importing works but
every call raises, loudly (`lib()` raises RuntimeError)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("MVB200_LIB", "libmvtools_b200.so"))  # MVB200_LIB: an experiment build (Makefile, V=)
PLUGIN_PATH = os.path.join(_HERE, "libmvtools.so")

UNSET = -(2 ** 31)
UNSET64 = -(2 ** 63)
MAX_LEVELS = 32
MAX_RADIUS = 6

i32, i64 = C.c_int32, C.c_int64


class Vector(C.Structure):
    _fields_ = [("x", i32), ("y", i32), ("sad", i64)]


class AnalysisData(C.Structure):
    _fields_ = [(n, i32) for n in (
        "nMagicKey", "nVersion", "nBlkSizeX", "nBlkSizeY", "nPel", "nLvCount", "nDeltaFrame", "isBackward",
        "nCPUFlags", "nMotionFlags", "nWidth", "nHeight", "nOverlapX", "nOverlapY", "nBlkX", "nBlkY",
        "bitsPerSample", "yRatioUV", "xRatioUV", "nHPadding", "nVPadding")]


class SuperParams(C.Structure):
    _fields_ = [(n, i32) for n in (
        "width", "height", "bits", "ssw", "ssh", "gray",
        "hpad", "vpad", "pel", "levels", "chroma", "sharp", "rfilter",
        "nplanes", "xRatioUV", "yRatioUV", "modeyuv", "super_width", "super_height")]


class LevelGeom(C.Structure):
    _fields_ = [(n, i32) for n in ("width", "height", "hpad", "vpad", "padded_w", "padded_h", "pel", "row0")]


class SuperGeom(C.Structure):
    _fields_ = [("nplanes", i32), ("levels", i32), ("plane_w", i32 * 3), ("plane_h", i32 * 3),
                ("lv", (LevelGeom * MAX_LEVELS) * 3)]


ANALYSE_ARGS = ("blksize", "blksizev", "levels", "search", "searchparam", "pelsearch", "isb", "lambda", "chroma",
                "delta", "truemotion", "lsad", "plevel", "global", "pnew", "pzero", "pglobal", "overlap", "overlapv",
                "divide", "badsad", "badrange", "opt", "meander", "trymany", "fields", "tff", "search_coarse", "dct")


class AnalyseArgs(C.Structure):
    _fields_ = [(n, i32) for n in ANALYSE_ARGS]


class AnalyseParams(C.Structure):
    _fields_ = [("ad", AnalysisData), ("divide", i32), ("ad_divided", AnalysisData)] + \
        [(n, i32) for n in ("search", "search_coarse", "nSearchParam", "nPelSearch", "nLambda", "lsad", "pnew",
                            "plevel", "global_", "pglobal", "pzero")] + \
        [("badSAD", i64)] + \
        [(n, i32) for n in ("badrange", "meander", "tryMany", "dctmode", "chroma", "fields", "tff", "tff_exists",
                            "super_hpad", "super_vpad", "super_pel", "super_modeyuv", "super_levels",
                            "ssw", "ssh", "gray", "num_frames")] + \
        [("nBlkX", i32 * MAX_LEVELS), ("nBlkY", i32 * MAX_LEVELS), ("vectors_size", i32)]


class DegrainArgs(C.Structure):
    _fields_ = [("radius", i32), ("thsad", i64), ("thsadc", i64), ("plane", i32), ("limit", i32), ("limitc", i32),
                ("thscd1", i64), ("thscd2", i32)]


class DegrainParams(C.Structure):
    _fields_ = [("radius", i32), ("vd", AnalysisData * (2 * MAX_RADIUS)), ("thSAD", i64 * 3), ("nLimit", i32 * 3),
                ("nSCD1", i64), ("nSCD2", i32), ("process", i32 * 3), ("YUVplanes", i32)] + \
        [(n, i32) for n in ("bits", "ssw", "ssh", "gray", "nplanes", "super_hpad", "super_vpad", "super_pel",
                            "super_modeyuv", "super_levels")] + \
        [(n, i32 * 3) for n in ("nWidth", "nHeight", "nOverlapX", "nOverlapY", "nBlkSizeX", "nBlkSizeY",
                                "nWidth_B", "nHeight_B")]


class CompensateArgs(C.Structure):
    _fields_ = [("scbehavior", i32), ("thsad", i64), ("fields", i32), ("time", C.c_double), ("thscd1", i64), ("thscd2", i32), ("tff", i32)]


class CompensateParams(C.Structure):
    _fields_ = [("vd", AnalysisData)] + [(n, i32) for n in ("scBehavior", "fields", "tff", "tff_exists", "time256")] + \
        [("thSAD", i64), ("nSCD1", i64), ("nSCD2", i32)] + \
        [(n, i32) for n in ("bits", "ssw", "ssh", "gray", "nplanes", "super_hpad", "super_vpad", "super_pel", "super_modeyuv", "super_levels")] + \
        [(n, i32 * 3) for n in ("nWidth", "nHeight", "nOverlapX", "nOverlapY", "nBlkSizeX", "nBlkSizeY", "nWidth_B", "nHeight_B")] + \
        [("num_frames", i32)]


RECALC_ARGS = ("smooth", "blksize", "blksizev", "search", "searchparam", "lambda", "chroma", "truemotion", "pnew", "overlap",
               "overlapv", "divide", "opt", "meander", "fields", "tff", "dct")


class RecalculateArgs(C.Structure):
    _fields_ = [("thsad", i64)] + [(n, i32) for n in RECALC_ARGS]


class RecalculateParams(C.Structure):
    _fields_ = [("ad", AnalysisData), ("ad_divided", AnalysisData), ("old", AnalysisData)] + \
        [(n, i32) for n in ("divide", "search", "nSearchParam", "nLambda", "pnew", "smooth", "meander", "dctmode", "chroma")] + \
        [("thSAD", i64)] + \
        [(n, i32) for n in ("fields", "tff", "tff_exists", "super_hpad", "super_vpad", "super_pel", "super_modeyuv", "super_levels",
                            "ssw", "ssh", "gray", "num_frames", "vectors_size")]


class ScdParams(C.Structure):
    _fields_ = [("vd", AnalysisData), ("nSCD1", i64), ("nSCD2", i32)]


class MaskArgs(C.Structure):
    _fields_ = [("ml", C.c_double), ("gamma", C.c_double), ("kind", i32), ("time", C.c_double), ("ysc", i32), ("thscd1", i64), ("thscd2", i32)]


class MaskParams(C.Structure):
    _fields_ = [("vd", AnalysisData)] + [(n, i32) for n in ("kind", "time256", "nSceneChangeValue")] + \
        [(n, C.c_float) for n in ("ml", "fGamma", "fMaskNormFactor", "fMaskNormFactor2", "fHalfGamma")] + \
        [("nSCD1", i64), ("nSCD2", i32), ("clip_gray", i32)] + \
        [(n, i32) for n in ("nWidth", "nHeight", "nWidthUV", "nHeightUV", "nWidthB", "nHeightB", "nWidthBUV", "nHeightBUV")]


class FlowArgs(C.Structure):
    _fields_ = [("time", C.c_double), ("mode", i32), ("fields", i32), ("thscd1", i64), ("thscd2", i32), ("tff", i32)]


class FlowParams(C.Structure):
    _fields_ = [("vd", AnalysisData)] + [(n, i32) for n in ("time256", "mode", "fields", "tff", "tff_exists")] + \
        [("nSCD1", i64), ("nSCD2", i32)] + \
        [(n, i32) for n in ("bits", "ssw", "ssh", "gray", "nplanes", "super_hpad", "super_vpad", "super_pel", "super_levels",
                            "nBlkXP", "nBlkYP", "nWidthP", "nHeightP", "nWidthPUV", "nHeightPUV", "nWidthUV", "nHeightUV",
                            "nHPaddingUV", "nVPaddingUV", "VPitchY", "VPitchUV", "num_frames")]


class FlowInterArgs(C.Structure):
    _fields_ = [("time", C.c_double), ("ml", C.c_double), ("blend", i32), ("thscd1", i64), ("thscd2", i32)]


class FlowInterParams(C.Structure):
    _fields_ = [("vd", AnalysisData), ("time256", i32), ("blend", i32), ("ml", C.c_double), ("nSCD1", i64), ("nSCD2", i32),
                ("mode_primary", i32), ("mode_fallback", i32)] + \
        [(n, i32) for n in ("bits", "ssw", "ssh", "gray", "nplanes", "super_hpad", "super_vpad", "super_pel", "super_levels",
                            "nBlkXP", "nBlkYP", "nWidthP", "nHeightP", "nWidthPUV", "nHeightPUV", "nWidthUV", "nHeightUV",
                            "nHPaddingUV", "nVPaddingUV", "VPitchY", "VPitchUV", "num_frames")]


class FlowFPSArgs(C.Structure):
    _fields_ = [("num", i64), ("den", i64), ("mask", i32), ("ml", C.c_double), ("blend", i32), ("thscd1", i64), ("thscd2", i32)]


class FlowFPSParams(C.Structure):
    _fields_ = [("inter", FlowInterParams), ("maskmode", i32), ("fa", i64), ("fb", i64), ("fps_num", i64), ("fps_den", i64),
                ("num_frames_in", i32), ("num_frames_out", i32)]


class FlowBlurArgs(C.Structure):
    _fields_ = [("blur", C.c_double), ("prec", i32), ("thscd1", i64), ("thscd2", i32)]


class FlowBlurParams(C.Structure):
    _fields_ = [("vd", AnalysisData), ("blur256", i32), ("prec", i32), ("nSCD1", i64), ("nSCD2", i32)] + \
        [(n, i32) for n in ("bits", "ssw", "ssh", "gray", "nplanes", "super_hpad", "super_vpad", "super_pel", "super_levels",
                            "nWidthUV", "nHeightUV", "nHPaddingUV", "nVPaddingUV", "num_frames")]


class BlockFPSArgs(C.Structure):
    _fields_ = [("num", i64), ("den", i64), ("mode", i32), ("ml", C.c_double), ("blend", i32), ("thscd1", i64), ("thscd2", i32)]


class BlockFPSParams(C.Structure):
    _fields_ = [("vd", AnalysisData), ("mode", i32), ("blend", i32), ("ml", C.c_double), ("nSCD1", i64), ("nSCD2", i32)] + \
        [(n, i32) for n in ("bits", "ssw", "ssh", "gray", "nplanes", "super_hpad", "super_vpad", "super_pel", "super_modeyuv", "super_levels",
                            "nBlkXP", "nBlkYP", "nWidthP", "nHeightP", "nWidthPUV", "nHeightPUV", "nWidthUV", "nHeightUV", "nPitchY", "nPitchUV")] + \
        [(n, i64) for n in ("fa", "fb", "fps_num", "fps_den")] + [("num_frames_in", i32), ("num_frames_out", i32)]


class FrameLayout(C.Structure):
    _fields_ = [("nplanes", i32), ("bytes_per_sample", i32), ("plane_w", i32 * 3), ("plane_h", i32 * 3),
                ("pitch", i64 * 3), ("offset", i64 * 3), ("frame_bytes", i64)]


class SearchPlan(C.Structure):
    _fields_ = [("kernel", i32), ("chains_per_sm", i32), ("chains_per_wave", i32), ("smem_per_chain", i32)]


class WindowPlan(C.Structure):
    _fields_ = [("frames", i32), ("contexts_per_device", i32), ("pinned_budget", i64)]


class Error(RuntimeError):
    pass


_LIB = None


def lib():
    """Loads libmvtools_b200.so (built in-tree by __graft_entry__.build()).  Raises if absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libmvtools_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        vp = C.c_void_p
        PP = C.POINTER
        L.mvb200_last_error.restype = C.c_char_p
        L.mvb200_version.restype = C.c_char_p
        L.mvb200_super_init.argtypes = [PP(SuperParams)]
        L.mvb200_super_geometry.argtypes = [PP(SuperParams), PP(SuperGeom)]
        L.mvb200_super_geometry.restype = None
        L.mvb200_analyse_init.argtypes = [PP(AnalyseArgs), PP(SuperParams), i32, PP(AnalyseParams)]
        L.mvb200_degrain_init.argtypes = [PP(DegrainArgs), PP(SuperParams), PP(AnalysisData), i32, i32, PP(DegrainParams)]
        L.mvb200_compensate_init.argtypes = [PP(CompensateArgs), PP(SuperParams), PP(AnalysisData), i32, i32, i32, PP(CompensateParams)]
        L.mvb200_recalculate_init.argtypes = [PP(RecalculateArgs), PP(SuperParams), PP(AnalysisData), i32, PP(RecalculateParams)]
        L.mvb200_compensate_frame.argtypes = [vp, PP(CompensateParams), PP(SuperParams), PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t), vp, i32,
                                              PP(vp), PP(C.c_ssize_t)]
        L.mvb200_recalculate_frame.argtypes = [vp, PP(RecalculateParams), PP(SuperParams), PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t), vp,
                                               i32, vp]
        L.mvb200_clip_layout.argtypes = [PP(SuperParams), PP(FrameLayout)]
        L.mvb200_clip_layout.restype = None
        L.mvb200_super_layout.argtypes = [PP(SuperParams), PP(FrameLayout)]
        L.mvb200_super_layout.restype = None
        L.mvb200_ctx_create.argtypes = [i32, PP(vp)]
        L.mvb200_ctx_destroy.argtypes = [vp]
        L.mvb200_ctx_destroy.restype = None
        L.mvb200_ctx_synchronize.argtypes = [vp]
        L.mvb200_ctx_stream.argtypes = [vp]
        L.mvb200_ctx_stream.restype = vp
        L.mvb200_ctx_launch_count.argtypes = [vp]
        L.mvb200_ctx_launch_count.restype = i64
        L.mvb200_search_plan_query.argtypes = [PP(AnalyseParams), PP(SuperParams), i32, PP(SearchPlan)]
        L.mvb200_window_plan_query.argtypes = [PP(AnalyseParams), PP(SuperParams), i32, i32, i64, i64, i32, i32, i64, i64, PP(WindowPlan)]
        L.mvb200_device_memory.argtypes = [i32]
        L.mvb200_device_memory.restype = i64
        L.mvb200_ctx_record.argtypes = [vp, PP(vp)]
        L.mvb200_event_wait.argtypes = [vp]
        L.mvb200_event_query.argtypes = [vp]
        L.mvb200_event_destroy.argtypes = [vp]
        L.mvb200_event_destroy.restype = None
        L.mvb200_super_batch.argtypes = [vp, PP(SuperParams), vp, vp, i32]
        L.mvb200_analyse_batch.argtypes = [vp, PP(AnalyseParams), PP(SuperParams), vp, i32, i32, vp, i64]
        L.mvb200_analyse_multi_batch.argtypes = [vp, PP(PP(AnalyseParams)), i32, PP(SuperParams), vp, i32, i32, PP(vp), i64]
        L.mvb200_degrain_batch.argtypes = [vp, PP(DegrainParams), PP(SuperParams), vp, vp, i32, i32, i32, PP(vp), i64, vp]
        L.mvb200_recalculate_batch.argtypes = [vp, PP(RecalculateParams), PP(SuperParams), vp, i32, i32, vp, i64, vp, i64]
        L.mvb200_compensate_batch.argtypes = [vp, PP(CompensateParams), PP(SuperParams), vp, i32, i32, vp, i64, vp]
        L.mvb200_degrain_clip_host_async.argtypes = [vp, PP(SuperParams), PP(PP(AnalyseParams)), i32, PP(DegrainParams), vp, vp, i32, i32, i32]
        L.mvb200_super_frame.argtypes = [vp, PP(SuperParams), PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t)]
        L.mvb200_super_frame_pel.argtypes = [vp, PP(SuperParams), PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t), i32, PP(vp), PP(C.c_ssize_t)]
        L.mvb200_super_pelclip_check.argtypes = [PP(SuperParams), i32, i32, i32, PP(i32), PP(i32)]
        L.mvb200_analyse_frame.argtypes = [vp, PP(AnalyseParams), PP(SuperParams), PP(vp), PP(C.c_ssize_t), PP(vp),
                                           PP(C.c_ssize_t), i32, vp]
        L.mvb200_degrain_frame.argtypes = [vp, PP(DegrainParams), PP(SuperParams), PP(vp), PP(C.c_ssize_t), PP(vp),
                                           PP(C.c_ssize_t), PP(vp), PP(vp), PP(C.c_ssize_t)]
        L.mvb200_scdetection_init.argtypes = [i64, i32, PP(AnalysisData), PP(ScdParams)]
        L.mvb200_mask_init.argtypes = [PP(MaskArgs), PP(AnalysisData), i32, i32, i32, i32, PP(MaskParams)]
        L.mvb200_flow_init.argtypes = [PP(FlowArgs), PP(SuperParams), PP(AnalysisData), i32, PP(FlowParams)]
        L.mvb200_simple_resize_tables.argtypes = [i32, i32, vp, vp]
        L.mvb200_simple_resize_tables.restype = None
        L.mvb200_mask_layout.argtypes = [PP(MaskParams), PP(FrameLayout)]
        L.mvb200_mask_layout.restype = None
        L.mvb200_finest_layout.argtypes = [PP(SuperParams), PP(FrameLayout)]
        L.mvb200_finest_layout.restype = None
        L.mvb200_scdetection_frame.argtypes = [vp, PP(ScdParams), vp, PP(i32)]
        L.mvb200_scdetection_batch.argtypes = [vp, PP(ScdParams), vp, i64, i32, PP(i32)]
        L.mvb200_mask_frame.argtypes = [vp, PP(MaskParams), vp, C.c_ssize_t, vp, PP(vp), PP(C.c_ssize_t)]
        L.mvb200_mask_batch.argtypes = [vp, PP(MaskParams), vp, i64, i64, vp, i64, i32, vp]
        L.mvb200_finest_frame.argtypes = [vp, PP(SuperParams), PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t)]
        L.mvb200_finest_batch.argtypes = [vp, PP(SuperParams), vp, i32, vp]
        L.mvb200_flow_frame.argtypes = [vp, PP(FlowParams), PP(SuperParams), PP(vp), PP(C.c_ssize_t), vp, i32, PP(vp), PP(C.c_ssize_t), PP(i32)]
        L.mvb200_flow_batch.argtypes = [vp, PP(FlowParams), PP(SuperParams), vp, vp, i32, i32, vp, i64, vp]
        L.mvb200_flowinter_init.argtypes = [PP(FlowInterArgs), PP(SuperParams), PP(AnalysisData), PP(AnalysisData), i32, PP(FlowInterParams)]
        L.mvb200_flowinter_frame.argtypes = [vp, PP(FlowInterParams), PP(SuperParams), i32, PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t), vp, vp, vp, vp, PP(vp), PP(C.c_ssize_t), PP(i32)]
        L.mvb200_blend_frame.argtypes = [vp, PP(SuperParams), i32, PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t)]
        L.mvb200_flowinter_batch.argtypes = [vp, PP(FlowInterParams), PP(SuperParams), vp, vp, i32, i32, vp, vp, i64, vp]
        L.mvb200_flowfps_init.argtypes = [PP(FlowFPSArgs), PP(SuperParams), PP(AnalysisData), PP(AnalysisData), i64, i64, i32, i32, i32, PP(FlowFPSParams)]
        L.mvb200_flowfps_position.argtypes = [PP(FlowFPSParams), i32, PP(i32), PP(i32)]
        L.mvb200_flowfps_position.restype = None
        L.mvb200_flowblur_init.argtypes = [PP(FlowBlurArgs), PP(SuperParams), PP(AnalysisData), PP(AnalysisData), i32, PP(FlowBlurParams)]
        L.mvb200_flowblur_frame.argtypes = [vp, PP(FlowBlurParams), PP(SuperParams), PP(vp), PP(C.c_ssize_t), vp, vp, PP(vp), PP(C.c_ssize_t), PP(i32)]
        L.mvb200_flowblur_batch.argtypes = [vp, PP(FlowBlurParams), PP(SuperParams), vp, vp, i32, i32, vp, vp, i64, vp]
        L.mvb200_blockfps_init.argtypes = [PP(BlockFPSArgs), PP(SuperParams), PP(AnalysisData), PP(AnalysisData), i64, i64, i32, i32, i32, PP(BlockFPSParams)]
        L.mvb200_blockfps_position.argtypes = [PP(BlockFPSParams), i32, PP(i32), PP(i32)]
        L.mvb200_blockfps_position.restype = None
        L.mvb200_blockfps_frame.argtypes = [vp, PP(BlockFPSParams), PP(SuperParams), i32, PP(vp), PP(C.c_ssize_t), PP(vp), PP(C.c_ssize_t), vp, vp, PP(vp), PP(C.c_ssize_t), PP(i32)]
        L.mvb200_analyse_frames.argtypes = [vp, PP(AnalyseParams), PP(SuperParams), i32, PP(vp), PP(C.c_ssize_t), i32, PP(i32), PP(i32), PP(i32), PP(vp)]
        L.mvb200_flowfps_clip_host.argtypes = [vp, PP(SuperParams), PP(AnalyseParams), PP(AnalyseParams), PP(FlowFPSParams), vp, i32, i32, i32, i32, vp, PP(i32)]
        L.mvb200_pinned_alloc.argtypes = [C.c_size_t]
        L.mvb200_pinned_alloc.restype = vp
        L.mvb200_pinned_free.argtypes = [vp]
        L.mvb200_pinned_free.restype = None
        L.mvb200_compensate_clip_host.argtypes = [vp, PP(SuperParams), PP(AnalyseParams), PP(CompensateParams), vp, i32, i32, i32, i32, vp]
        L.mvb200_flow_clip_host.argtypes = [vp, PP(SuperParams), PP(AnalyseParams), PP(FlowParams), vp, i32, i32, i32, i32, vp, PP(i32)]
        L.mvb200_flowinter_clip_host.argtypes = [vp, PP(SuperParams), PP(AnalyseParams), PP(AnalyseParams), PP(FlowInterParams), vp, i32, i32, i32, i32, vp, PP(i32)]
        L.mvb200_blockfps_clip_host.argtypes = [vp, PP(SuperParams), PP(AnalyseParams), PP(AnalyseParams), PP(BlockFPSParams), vp, i32, i32, i32, i32, vp, PP(i32)]
        L.mvb200_flowblur_clip_host.argtypes = [vp, PP(SuperParams), PP(AnalyseParams), PP(AnalyseParams), PP(FlowBlurParams), vp, i32, i32, i32, i32, vp]
        _LIB = L
    return _LIB


def super_pelclip_check(sp, pel_width, pel_height, same_format=True):
    """-> (use_pelclip, pel_padded); raises Error with the reference's message"""
    use, padded = C.c_int32(0), C.c_int32(0)
    _check(lib().mvb200_super_pelclip_check(C.byref(sp), int(same_format), pel_width, pel_height, C.byref(use), C.byref(padded)))
    return bool(use.value), bool(padded.value)


def _check(rc):
    if rc:
        raise Error(lib().mvb200_last_error().decode())


# ----------------------------------------------------------------- host logic

def super_params(width, height, bits=8, ssw=1, ssh=1, gray=False, **kw):
    p = SuperParams()
    p.width, p.height, p.bits, p.ssw, p.ssh, p.gray = width, height, bits, ssw, ssh, int(gray)
    for n in ("hpad", "vpad", "pel", "levels", "chroma", "sharp", "rfilter"):
        setattr(p, n, int(kw.pop(n)) if n in kw and kw[n] is not None else UNSET)
    if kw:
        raise TypeError("Super: unknown argument(s) %s" % sorted(kw))
    _check(lib().mvb200_super_init(C.byref(p)))
    return p


def super_geometry(sp):
    g = SuperGeom()
    lib().mvb200_super_geometry(C.byref(sp), C.byref(g))
    return g


def analyse_params(sp, num_frames, **kw):
    a = AnalyseArgs()
    for n in ANALYSE_ARGS:
        v = kw.pop(n, None)
        setattr(a, n, UNSET if v is None else int(v))
    if kw:
        raise TypeError("Analyse: unknown argument(s) %s" % sorted(kw))
    out = AnalyseParams()
    _check(lib().mvb200_analyse_init(C.byref(a), C.byref(sp), num_frames, C.byref(out)))
    return out


def degrain_params(sp, vds, clip_w, clip_h, radius, **kw):
    a = DegrainArgs()
    a.radius = radius
    for n, unset in (("thsad", UNSET64), ("thsadc", UNSET64), ("plane", UNSET), ("limit", UNSET), ("limitc", UNSET),
                     ("thscd1", UNSET64), ("thscd2", UNSET)):
        v = kw.pop(n, None)
        setattr(a, n, unset if v is None else int(v))
    if kw:
        raise TypeError("Degrain: unknown argument(s) %s" % sorted(kw))
    arr = (AnalysisData * len(vds))(*vds)
    out = DegrainParams()
    _check(lib().mvb200_degrain_init(C.byref(a), C.byref(sp), arr, clip_w, clip_h, C.byref(out)))
    return out


def compensate_params(sp, vd, clip_w, clip_h, num_frames, **kw):
    a = CompensateArgs()
    a.scbehavior = UNSET if kw.get("scbehavior") is None else int(kw.pop("scbehavior"))
    kw.pop("scbehavior", None)
    a.thsad = UNSET64 if kw.get("thsad") is None else int(kw.pop("thsad"))
    kw.pop("thsad", None)
    a.fields = UNSET if kw.get("fields") is None else int(kw.pop("fields"))
    kw.pop("fields", None)
    t = kw.pop("time", None)
    a.time = float("nan") if t is None else float(t)
    a.thscd1 = UNSET64 if kw.get("thscd1") is None else int(kw.pop("thscd1"))
    kw.pop("thscd1", None)
    a.thscd2 = UNSET if kw.get("thscd2") is None else int(kw.pop("thscd2"))
    kw.pop("thscd2", None)
    a.tff = UNSET if kw.get("tff") is None else int(kw.pop("tff"))
    kw.pop("tff", None)
    if kw:
        raise TypeError("Compensate: unknown argument(s) %s" % sorted(kw))
    out = CompensateParams()
    _check(lib().mvb200_compensate_init(C.byref(a), C.byref(sp), C.byref(vd), clip_w, clip_h, num_frames, C.byref(out)))
    return out


def recalculate_params(sp, old_vd, num_frames, **kw):
    a = RecalculateArgs()
    v = kw.pop("thsad", None)
    a.thsad = UNSET64 if v is None else int(v)
    for n in RECALC_ARGS:
        v = kw.pop(n, None)
        setattr(a, n, UNSET if v is None else int(v))
    if kw:
        raise TypeError("Recalculate: unknown argument(s) %s" % sorted(kw))
    out = RecalculateParams()
    _check(lib().mvb200_recalculate_init(C.byref(a), C.byref(sp), C.byref(old_vd), num_frames, C.byref(out)))
    return out


def _opt(kw, name, unset, conv=int):
    v = kw.pop(name, None)
    return unset if v is None else conv(v)


def scdetection_params(vd, thscd1=None, thscd2=None):
    out = ScdParams()
    _check(lib().mvb200_scdetection_init(UNSET64 if thscd1 is None else int(thscd1), UNSET if thscd2 is None else int(thscd2), C.byref(vd), C.byref(out)))
    return out


def mask_params(vd, clip_bits=8, clip_ssw=1, clip_ssh=1, clip_gray=False, **kw):
    a = MaskArgs()
    nan = float("nan")
    a.ml, a.gamma, a.time = _opt(kw, "ml", nan, float), _opt(kw, "gamma", nan, float), _opt(kw, "time", nan, float)
    a.kind, a.ysc, a.thscd2 = _opt(kw, "kind", UNSET), _opt(kw, "ysc", UNSET), _opt(kw, "thscd2", UNSET)
    a.thscd1 = _opt(kw, "thscd1", UNSET64)
    if kw:
        raise TypeError("Mask: unknown argument(s) %s" % sorted(kw))
    out = MaskParams()
    _check(lib().mvb200_mask_init(C.byref(a), C.byref(vd), clip_bits, clip_ssw, clip_ssh, int(clip_gray), C.byref(out)))
    return out


def flow_params(sp, vd, num_frames, **kw):
    a = FlowArgs()
    a.time = _opt(kw, "time", float("nan"), float)
    a.mode, a.fields, a.thscd2, a.tff = _opt(kw, "mode", UNSET), _opt(kw, "fields", UNSET), _opt(kw, "thscd2", UNSET), _opt(kw, "tff", UNSET)
    a.thscd1 = _opt(kw, "thscd1", UNSET64)
    if kw:
        raise TypeError("Flow: unknown argument(s) %s" % sorted(kw))
    out = FlowParams()
    _check(lib().mvb200_flow_init(C.byref(a), C.byref(sp), C.byref(vd), num_frames, C.byref(out)))
    return out


def flowinter_params(sp, vd_bw, vd_fw, num_frames, **kw):
    a = FlowInterArgs()
    a.time, a.ml = _opt(kw, "time", float("nan"), float), _opt(kw, "ml", float("nan"), float)
    a.blend, a.thscd2 = _opt(kw, "blend", UNSET), _opt(kw, "thscd2", UNSET)
    a.thscd1 = _opt(kw, "thscd1", UNSET64)
    if kw:
        raise TypeError("FlowInter: unknown argument(s) %s" % sorted(kw))
    out = FlowInterParams()
    _check(lib().mvb200_flowinter_init(C.byref(a), C.byref(sp), C.byref(vd_bw), C.byref(vd_fw), num_frames, C.byref(out)))
    return out


def flowfps_params(sp, vd_bw, vd_fw, num_frames, clip_w, clip_h, fps_num=25, fps_den=1, **kw):
    a = FlowFPSArgs()
    a.num, a.den, a.thscd1 = _opt(kw, "num", UNSET64), _opt(kw, "den", UNSET64), _opt(kw, "thscd1", UNSET64)
    a.mask, a.blend, a.thscd2 = _opt(kw, "mask", UNSET), _opt(kw, "blend", UNSET), _opt(kw, "thscd2", UNSET)
    a.ml = _opt(kw, "ml", float("nan"), float)
    if kw:
        raise TypeError("FlowFPS: unknown argument(s) %s" % sorted(kw))
    out = FlowFPSParams()
    _check(lib().mvb200_flowfps_init(C.byref(a), C.byref(sp), C.byref(vd_bw), C.byref(vd_fw), fps_num, fps_den, clip_w, clip_h, num_frames, C.byref(out)))
    return out


def flowfps_position(p, n):
    nleft, t = i32(0), i32(0)
    lib().mvb200_flowfps_position(C.byref(p), n, C.byref(nleft), C.byref(t))
    return nleft.value, t.value


def flowblur_params(sp, vd_bw, vd_fw, num_frames, **kw):
    a = FlowBlurArgs()
    a.blur = _opt(kw, "blur", float("nan"), float)
    a.prec, a.thscd2 = _opt(kw, "prec", UNSET), _opt(kw, "thscd2", UNSET)
    a.thscd1 = _opt(kw, "thscd1", UNSET64)
    if kw:
        raise TypeError("FlowBlur: unknown argument(s) %s" % sorted(kw))
    out = FlowBlurParams()
    _check(lib().mvb200_flowblur_init(C.byref(a), C.byref(sp), C.byref(vd_bw), C.byref(vd_fw), num_frames, C.byref(out)))
    return out


def blockfps_params(sp, vd_bw, vd_fw, num_frames, clip_w, clip_h, fps_num=25, fps_den=1, **kw):
    a = BlockFPSArgs()
    a.num, a.den, a.thscd1 = _opt(kw, "num", UNSET64), _opt(kw, "den", UNSET64), _opt(kw, "thscd1", UNSET64)
    a.mode, a.blend, a.thscd2 = _opt(kw, "mode", UNSET), _opt(kw, "blend", UNSET), _opt(kw, "thscd2", UNSET)
    a.ml = _opt(kw, "ml", float("nan"), float)
    if kw:
        raise TypeError("BlockFPS: unknown argument(s) %s" % sorted(kw))
    out = BlockFPSParams()
    _check(lib().mvb200_blockfps_init(C.byref(a), C.byref(sp), C.byref(vd_bw), C.byref(vd_fw), fps_num, fps_den, clip_w, clip_h, num_frames, C.byref(out)))
    return out


def blockfps_position(p, n):
    nleft, t = i32(0), i32(0)
    lib().mvb200_blockfps_position(C.byref(p), n, C.byref(nleft), C.byref(t))
    return nleft.value, t.value


def simple_resize_tables(out, in_):
    off = np.zeros(out, dtype=np.int32)
    wt = np.zeros(out, dtype=np.int32)
    lib().mvb200_simple_resize_tables(out, in_, off.ctypes.data, wt.ctypes.data)
    return off, wt


def mask_layout(mp):
    l = FrameLayout()
    lib().mvb200_mask_layout(C.byref(mp), C.byref(l))
    return l


def finest_layout(sp):
    l = FrameLayout()
    lib().mvb200_finest_layout(C.byref(sp), C.byref(l))
    return l


def clip_layout(sp):
    l = FrameLayout()
    lib().mvb200_clip_layout(C.byref(sp), C.byref(l))
    return l


def super_layout(sp):
    l = FrameLayout()
    lib().mvb200_super_layout(C.byref(sp), C.byref(l))
    return l


# ------------------------------------------------------ plane <-> pointer glue

def plane_ptrs(planes):
    """[np.ndarray]*n -> (void*[3], ssize_t[3]); arrays must be C-contiguous rows."""
    ptrs = (C.c_void_p * 3)()
    strides = (C.c_ssize_t * 3)()
    for i, a in enumerate(planes):
        if a is None:
            continue
        assert a.strides[-1] == a.itemsize
        ptrs[i] = a.ctypes.data
        strides[i] = a.strides[0]
    return ptrs, strides


def alloc_planes(widths, heights, bits, fill=None):
    dt = np.uint8 if bits <= 8 else np.uint16
    out = []
    for w, h in zip(widths, heights):
        a = np.empty((h, w), dtype=dt)
        if fill is not None:
            a[...] = fill
        out.append(a)
    return out


class Context:
    """One CUDA device context (stream + scratch) of the engine."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().mvb200_ctx_create(device, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().mvb200_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def synchronize(self):
        _check(lib().mvb200_ctx_synchronize(self._h))

    def stream(self):
        return lib().mvb200_ctx_stream(self._h)

    def launch_count(self):
        return lib().mvb200_ctx_launch_count(self._h)

    # ---- host-buffer per-frame calls (what the VapourSynth plugin makes)
    def super_frame(self, sp, src_planes, pel_planes=None, pel_padded=False):
        if pel_planes is not None:
            return self.super_frame_pel(sp, src_planes, pel_planes, pel_padded)
        g = super_geometry(sp)
        dst = alloc_planes(list(g.plane_w)[: sp.nplanes], list(g.plane_h)[: sp.nplanes], sp.bits, fill=0xA5)
        s, ss = plane_ptrs(src_planes)
        d, ds = plane_ptrs(dst)
        _check(lib().mvb200_super_frame(self._h, C.byref(sp), s, ss, d, ds))
        return dst

    def super_frame_pel(self, sp, src_planes, pel_planes, pel_padded):
        g = super_geometry(sp)
        dst = alloc_planes(list(g.plane_w)[: sp.nplanes], list(g.plane_h)[: sp.nplanes], sp.bits, fill=0xA5)
        s, ss = plane_ptrs(src_planes)
        q, qs = plane_ptrs(pel_planes)
        d, ds = plane_ptrs(dst)
        _check(lib().mvb200_super_frame_pel(self._h, C.byref(sp), s, ss, q, qs, int(pel_padded), d, ds))
        return dst

    def analyse_frame(self, ap, sp, src_super, ref_super, field_shift=0):
        out = np.empty(ap.vectors_size, dtype=np.uint8)
        s, ss = plane_ptrs(src_super)
        if ref_super is None:
            r, rs = None, None
        else:
            r, rs = plane_ptrs(ref_super)
        _check(lib().mvb200_analyse_frame(self._h, C.byref(ap), C.byref(sp), s, ss, r, rs, field_shift, out.ctypes.data))
        return out

    def recalculate_frame(self, rp, sp, src_super, ref_super, old_vectors, field_shift=0):
        out = np.empty(rp.vectors_size, dtype=np.uint8)
        s, ss = plane_ptrs(src_super)
        r, rs = (None, None) if ref_super is None else plane_ptrs(ref_super)
        _check(lib().mvb200_recalculate_frame(self._h, C.byref(rp), C.byref(sp), s, ss, r, rs, old_vectors.ctypes.data, field_shift, out.ctypes.data))
        return out

    def compensate_frame(self, cp, sp, src_super, ref_super, vectors, field_shift=0):
        dst = alloc_planes(list(cp.nWidth)[: cp.nplanes], list(cp.nHeight)[: cp.nplanes], cp.bits, fill=0xA5)
        s, ss = plane_ptrs(src_super)
        r, rs = (None, None) if ref_super is None else plane_ptrs(ref_super)
        d, ds = plane_ptrs(dst)
        _check(lib().mvb200_compensate_frame(self._h, C.byref(cp), C.byref(sp), s, ss, r, rs, vectors.ctypes.data, field_shift, d, ds))
        return dst

    def degrain_frame(self, dp, sp, src_planes, ref_supers, vectors):
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
        dst = alloc_planes([a.shape[1] for a in src_planes], [a.shape[0] for a in src_planes], dp.bits, fill=0xA5)
        s, ss = plane_ptrs(src_planes)
        d, ds = plane_ptrs(dst)
        _check(lib().mvb200_degrain_frame(self._h, C.byref(dp), C.byref(sp), s, ss, refs, rstr, vecs, d, ds))
        return dst

    # ---- consumers of the vector clips
    def scdetection_frame(self, scp, vectors):
        out = i32(0)
        _check(lib().mvb200_scdetection_frame(self._h, C.byref(scp), vectors.ctypes.data, C.byref(out)))
        return out.value

    def mask_frame(self, mp, src_luma, vectors):
        dst = alloc_planes([mp.nWidth, mp.nWidthUV, mp.nWidthUV], [mp.nHeight, mp.nHeightUV, mp.nHeightUV], 8, fill=0xA5)
        d, ds = plane_ptrs(dst)
        sl = None if src_luma is None else src_luma.ctypes.data
        _check(lib().mvb200_mask_frame(self._h, C.byref(mp), sl, 0 if src_luma is None else src_luma.strides[0], vectors.ctypes.data, d, ds))
        return dst

    def finest_frame(self, sp, super_planes):
        l = finest_layout(sp)
        dst = alloc_planes(list(l.plane_w)[: sp.nplanes], list(l.plane_h)[: sp.nplanes], sp.bits, fill=0xA5)
        s, ss = plane_ptrs(super_planes)
        d, ds = plane_ptrs(dst)
        _check(lib().mvb200_finest_frame(self._h, C.byref(sp), s, ss, d, ds))
        return dst

    def flow_frame(self, fp, sp, ref_super, vectors, field_shift=0):
        """-> planes, or None when the vectors are not usable (the caller keeps the source frame)"""
        w = [fp.vd.nWidth, fp.nWidthUV, fp.nWidthUV][: fp.nplanes]
        h = [fp.vd.nHeight, fp.nHeightUV, fp.nHeightUV][: fp.nplanes]
        dst = alloc_planes(w, h, fp.bits, fill=0xA5)
        r, rs = (None, None) if ref_super is None else plane_ptrs(ref_super)
        d, ds = plane_ptrs(dst)
        usable = i32(0)
        _check(lib().mvb200_flow_frame(self._h, C.byref(fp), C.byref(sp), r, rs, vectors.ctypes.data, field_shift, d, ds, C.byref(usable)))
        return dst if usable.value else None

    def flowinter_frame(self, p, sp, time256, left_super, right_super, vec_b, vec_f, vec_bb=None, vec_ff=None):
        """-> planes, or None when the vectors are not usable (the caller blends or keeps the source frame)"""
        w = [p.vd.nWidth, p.nWidthUV, p.nWidthUV][: p.nplanes]
        h = [p.vd.nHeight, p.nHeightUV, p.nHeightUV][: p.nplanes]
        dst = alloc_planes(w, h, p.bits, fill=0xA5)
        l, ls = plane_ptrs(left_super)
        r, rs = plane_ptrs(right_super)
        d, ds = plane_ptrs(dst)
        usable = i32(0)
        ptr = lambda v: None if v is None else v.ctypes.data
        _check(lib().mvb200_flowinter_frame(self._h, C.byref(p), C.byref(sp), time256, l, ls, r, rs, ptr(vec_b), ptr(vec_f), ptr(vec_bb), ptr(vec_ff), d, ds, C.byref(usable)))
        return dst if usable.value else None

    def blend_frame(self, sp, time256, src, ref):
        dst = alloc_planes([a.shape[1] for a in src], [a.shape[0] for a in src], sp.bits, fill=0xA5)
        s, ss = plane_ptrs(src)
        r, rs = plane_ptrs(ref)
        d, ds = plane_ptrs(dst)
        _check(lib().mvb200_blend_frame(self._h, C.byref(sp), time256, s, ss, r, rs, d, ds))
        return dst

    def flowblur_frame(self, p, sp, super_planes, vec_b, vec_f):
        """-> planes, or None when the vectors are not usable (the caller keeps the source frame)"""
        w = [p.vd.nWidth, p.nWidthUV, p.nWidthUV][: p.nplanes]
        h = [p.vd.nHeight, p.nHeightUV, p.nHeightUV][: p.nplanes]
        dst = alloc_planes(w, h, p.bits, fill=0xA5)
        s, ss = plane_ptrs(super_planes)
        d, ds = plane_ptrs(dst)
        usable = i32(0)
        ptr = lambda v: None if v is None else v.ctypes.data
        _check(lib().mvb200_flowblur_frame(self._h, C.byref(p), C.byref(sp), s, ss, ptr(vec_b), ptr(vec_f), d, ds, C.byref(usable)))
        return dst if usable.value else None

    def blockfps_frame(self, p, sp, time256, left_super, right_super, vec_b, vec_f):
        """-> planes, or None when the vectors are not usable"""
        w = [p.vd.nWidth, p.nWidthUV, p.nWidthUV][: p.nplanes]
        h = [p.vd.nHeight, p.nHeightUV, p.nHeightUV][: p.nplanes]
        dst = alloc_planes(w, h, p.bits, fill=0xA5)
        l, ls = plane_ptrs(left_super)
        r, rs = plane_ptrs(right_super)
        d, ds = plane_ptrs(dst)
        usable = i32(0)
        _check(lib().mvb200_blockfps_frame(self._h, C.byref(p), C.byref(sp), time256, l, ls, r, rs, vec_b.ctypes.data, vec_f.ctypes.data, d, ds, C.byref(usable)))
        return dst if usable.value else None

    def analyse_frames(self, ap, sp, super_frames, jobs):
        """One engine call for several frames of one mv.Analyse instance: super_frames = list of plane lists (uploaded
        once), jobs = [(src index, ref index or None, field shift)] -> list of blobs"""
        nf, nj = len(super_frames), len(jobs)
        planes = (C.c_void_p * (3 * nf))()
        strides = (C.c_ssize_t * (3 * nf))()
        for i, fr in enumerate(super_frames):
            for p, a in enumerate(fr):
                planes[3 * i + p] = a.ctypes.data
                strides[3 * i + p] = a.strides[0]
        src = (i32 * nj)(*[j[0] for j in jobs])
        ref = (i32 * nj)(*[-1 if j[1] is None else j[1] for j in jobs])
        fs = (i32 * nj)(*[j[2] for j in jobs])
        outs = [np.empty(ap.vectors_size, dtype=np.uint8) for _ in jobs]
        outp = (C.c_void_p * nj)(*[o.ctypes.data for o in outs])
        _check(lib().mvb200_analyse_frames(self._h, C.byref(ap), C.byref(sp), nf, planes, strides, nj, src, ref, fs, outp))
        return outs
