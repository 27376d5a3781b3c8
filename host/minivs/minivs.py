"""ctypes binding of the mini VapourSynth host (libminivs.so) -- a frame-server stand-in (this image has no VapourSynth):
it loads a VapourSynth API-4 plugin (this repo's libmvtools.so, or the compiled reference) and drives it the way vspipe would.

Gives tests a tiny imitation of the VapourSynth Python API:
    core = Core(); mv = core.load(path_to_plugin)
    clip = core.source(frames, bits=8, ssw=1, ssh=1)
    sup  = mv.Super(clip, pel=2)
    vec  = mv.Analyse(sup, isb=1, blksize=8)
    f = vec.get_frame(3); f.planes(); f.prop_data("MVTools_vectors")
so the parity tests read like a .vpy script and run unchanged against the reference
build (oracle/_ref/libmvtools_ref.so) and this repo's drop-in plugin."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

cfGray, cfYUV = 1, 3


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libminivs.so")
        L = C.CDLL(os.path.abspath(path), mode=C.RTLD_GLOBAL)
        vp, cp, i32, i64 = C.c_void_p, C.c_char_p, C.c_int, C.c_int64
        sig = {
            "mvs_last_error": (cp, []),
            "mvs_core_create": (vp, []),
            "mvs_core_free": (None, [vp]),
            "mvs_load_plugin": (vp, [vp, cp]),
            "mvs_plugin_info": (cp, [vp, i32]),
            "mvs_plugin_num_functions": (i32, [vp]),
            "mvs_plugin_function_name": (cp, [vp, i32]),
            "mvs_plugin_function_args": (cp, [vp, cp]),
            "mvs_map_create": (vp, []),
            "mvs_map_free": (None, [vp]),
            "mvs_map_set_int": (None, [vp, cp, i64]),
            "mvs_map_set_float": (None, [vp, cp, C.c_double]),
            "mvs_map_set_node": (None, [vp, cp, vp]),
            "mvs_map_error": (cp, [vp]),
            "mvs_map_get_node": (vp, [vp, cp]),
            "mvs_invoke": (vp, [vp, cp, vp]),
            "mvs_source_create": (vp, [vp, i32, i32, i32, i32, i32, i32, i32]),
            "mvs_source_set_plane": (None, [vp, i32, i32, vp, i64]),
            "mvs_source_alias": (None, [vp, i32, i32]),
            "mvs_source_set_prop_int": (None, [vp, i32, cp, i64]),
            "mvs_source_set_delay": (None, [vp, i32, i32]),
            "mvs_node_free": (None, [vp]),
            "mvs_node_set_cache": (None, [vp, i32]),
            "mvs_node_computed": (C.c_long, [vp]),
            "mvs_node_info": (None, [vp, C.POINTER(i32)]),
            "mvs_get_frame": (vp, [vp, i32]),
            "mvs_frame_free": (None, [vp]),
            "mvs_frame_width": (i32, [vp, i32]),
            "mvs_frame_height": (i32, [vp, i32]),
            "mvs_frame_stride": (i64, [vp, i32]),
            "mvs_frame_plane": (vp, [vp, i32]),
            "mvs_frame_prop_int": (i32, [vp, cp, C.POINTER(i64)]),
            "mvs_frame_prop_data_size": (i32, [vp, cp]),
            "mvs_frame_prop_data": (vp, [vp, cp]),
            "mvs_render": (C.c_double, [vp, i32, i32, i32]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


class Error(RuntimeError):
    pass


class Frame:
    def __init__(self, ptr, info):
        self._p = ptr
        self._info = info

    def __del__(self):
        if self._p:
            lib().mvs_frame_free(self._p)
            self._p = None

    def plane(self, p):
        L = lib()
        w, h = L.mvs_frame_width(self._p, p), L.mvs_frame_height(self._p, p)
        stride = L.mvs_frame_stride(self._p, p)
        bps = self._info[5]
        buf = (C.c_uint8 * (stride * h)).from_address(L.mvs_frame_plane(self._p, p))
        a = np.frombuffer(buf, dtype=np.uint8).reshape(h, stride)[:, : w * bps]
        a = np.array(a, copy=True, order="C")  # (always a copy: the frame may be freed before the array is read)
        return a.view(np.uint16) if bps == 2 else a

    def planes(self):
        return [self.plane(p) for p in range(self._info[8])]

    def prop_int(self, key):
        v = C.c_int64(0)
        if not lib().mvs_frame_prop_int(self._p, key.encode(), C.byref(v)):
            return None
        return v.value

    def prop_data(self, key):
        L = lib()
        n = L.mvs_frame_prop_data_size(self._p, key.encode())
        if n < 0:
            return None
        return C.string_at(L.mvs_frame_prop_data(self._p, key.encode()), n)


class Node:
    def __init__(self, ptr):
        self._p = ptr
        info = (C.c_int * 9)()
        lib().mvs_node_info(ptr, info)
        self.info = list(info)
        self.width, self.height, self.num_frames = self.info[0:3]
        self.bits = self.info[4]

    def __del__(self):
        if self._p:
            lib().mvs_node_free(self._p)
            self._p = None

    def get_frame(self, n):
        f = lib().mvs_get_frame(self._p, n)
        if not f:
            raise Error(lib().mvs_last_error().decode())
        return Frame(f, self.info)

    def set_cache(self, n):
        lib().mvs_node_set_cache(self._p, n)

    def computed(self):
        return lib().mvs_node_computed(self._p)

    def render(self, n0, n1, threads):
        t = lib().mvs_render(self._p, n0, n1, threads)
        if t < 0:
            raise Error(lib().mvs_last_error().decode())
        return t


class Plugin:
    def __init__(self, ptr):
        self._p = ptr

    @property
    def identifier(self):
        return lib().mvs_plugin_info(self._p, 0).decode()

    @property
    def namespace(self):
        return lib().mvs_plugin_info(self._p, 1).decode()

    def functions(self):
        L = lib()
        out = {}
        for i in range(L.mvs_plugin_num_functions(self._p)):
            name = L.mvs_plugin_function_name(self._p, i)
            out[name.decode()] = L.mvs_plugin_function_args(self._p, name).decode()
        return out

    def invoke(self, name, *pos, **kw):
        L = lib()
        sig = self.functions()[name]
        names = [a.split(":")[0] for a in sig.split(";") if a]
        args = dict(zip(names, pos))
        args.update(kw)
        m = L.mvs_map_create()
        keep = []
        for k, v in args.items():
            if v is None:
                continue
            if isinstance(v, Node):
                L.mvs_map_set_node(m, k.encode(), v._p)
                keep.append(v)
            elif isinstance(v, float):
                L.mvs_map_set_float(m, k.encode(), v)
            else:
                L.mvs_map_set_int(m, k.encode(), int(v))
        out = L.mvs_invoke(self._p, name.encode(), m)
        L.mvs_map_free(m)
        err = L.mvs_map_error(out)
        if err is not None:
            msg = err.decode()
            L.mvs_map_free(out)
            raise Error(msg)
        node = L.mvs_map_get_node(out, b"clip")
        L.mvs_map_free(out)
        return Node(node)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return lambda *a, **k: self.invoke(name, *a, **k)


class Core:
    def __init__(self):
        self._p = lib().mvs_core_create()

    def load(self, path):
        p = lib().mvs_load_plugin(self._p, os.path.abspath(path).encode())
        if not p:
            raise Error(lib().mvs_last_error().decode())
        return Plugin(p)

    def source(self, frames, bits=8, ssw=1, ssh=1, props=None, delay_ms=None):
        """frames: list of [Y,U,V] (or [Y]) numpy arrays; props: optional list (one dict of integer frame properties per frame);
        delay_ms: optional {frame: milliseconds} -- fetching those frames is slow (to order concurrent requests in a test).
        A frame that IS (same list object) an earlier frame of the list shares that frame's memory (cyclic clips)."""
        L = lib()
        seen = {}
        nplanes = len(frames[0])
        h, w = frames[0][0].shape
        node = L.mvs_source_create(self._p, cfGray if nplanes == 1 else cfYUV, bits,
                                   ssw if nplanes > 1 else 0, ssh if nplanes > 1 else 0, w, h, len(frames))
        if not node:
            raise Error(L.mvs_last_error().decode())
        for n, fr in enumerate(frames):
            if id(fr) in seen and not props:
                L.mvs_source_alias(node, n, seen[id(fr)])
                continue
            seen[id(fr)] = n
            for p, a in enumerate(fr):
                a = np.ascontiguousarray(a)
                L.mvs_source_set_plane(node, n, p, a.ctypes.data, a.strides[0])
            for k, v in (props[n] if props else {}).items():
                L.mvs_source_set_prop_int(node, n, k.encode(), int(v))
        for n, ms in (delay_ms or {}).items():
            L.mvs_source_set_delay(node, int(n), int(ms))
        return Node(node)


REF_PLUGIN = os.path.abspath(os.path.join(_HERE, "..", "..", "oracle", "_ref", "libmvtools_ref.so"))
