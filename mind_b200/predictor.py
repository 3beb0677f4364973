"""Drop-in scene predictor: the class MIND imports by string.

    net_cfg["network"] = "mind_b200.predictor:ScenePredNetB200"

mirrors the call surface MINDPlanner and ScenarioTreeGenerator use on the reference's
ScenePredNet (reference planners/mind/planner.py:42-49, planners/mind/scenario_tree.py:69-71,
planners/mind/networks/network.py:559-606): __init__(cfg, device), load_state_dict, to, eval,
pre_process(data) -> 7-tuple, __call__(data_in) -> (res_cls, res_reg, res_aux).

All arithmetic runs in libmind_b200.so (hand-written sm_100a CUDA); torch is used for device
memory, streams and host<->device copies only.  No CPU fallback: a non-CUDA device raises.
"""
import ctypes as C
import math
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import lib as _lib

_EXPECTED_CFG = dict(in_actor=14, d_actor=128, n_fpn_scale=4, in_lane=16, d_lane=128, d_rpe_in=5, d_rpe=128,
                     d_embed=128, n_scene_layer=6, n_scene_head=8, g_num_modes=6, g_pred_len=60,
                     param_out="bezier", update_edge=True)


def _bezier_bases(n_order=7, n_step=60):
    """Same construction as the reference (network.py:449-464): float64 numpy, cast to fp32."""
    ts = np.linspace(0.0, 1.0, n_step, endpoint=True)
    T = np.array([math.comb(n_order, i) * (1.0 - ts) ** (n_order - i) * ts ** i for i in range(n_order + 1)]).T
    Tp = np.array([n_order * math.comb(n_order - 1, i) * (1.0 - ts) ** (n_order - 1 - i) * ts ** i
                   for i in range(n_order)]).T
    return np.ascontiguousarray(T, dtype=np.float32), np.ascontiguousarray(Tp, dtype=np.float32)


def _pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def pack_layout(B: int, A: int):
    """(offset, numel) of cls [B,6], reg [A,6,60,5], vel [A,6,60,2] inside forward_packed's output buffer `last_pack`."""
    n_cls, n_reg = _pad64(B * 6), _pad64(A * 1800)
    return (0, B * 6), (n_cls, A * 1800), (n_cls + n_reg, A * 720)


def _to_device(data, device):
    """Recursive transfer, same contract as the reference's gpu() (planners/mind/utils.py:9-20)."""
    if isinstance(data, (list, tuple)):
        return [_to_device(x, device) for x in data]
    if isinstance(data, dict):
        return {k: _to_device(v, device) for k, v in data.items()}
    if isinstance(data, torch.Tensor):
        return data.contiguous().to(device, non_blocking=True)
    return data


class _PackedRPE(list):
    """data['RPE'] after pre_process: the usual list of {'scene': [5,M,M] device tensor, 'scene_mask': None}, whose
    'scene' tensors are views into ONE device buffer filled by a single mind_upload_packed call.  forward_packed
    reads the device pointers from `ptrs` instead of touching the per-scene tensors."""
    base = None      # the packed device buffer (uint8)
    ptrs = None      # ctypes array of per-scene device pointers
    shapes = None    # per-scene (5, M, M)


class _LazyRPE(dict):
    """One entry of data['RPE'] when only the anchors travelled: 'scene' is evaluated on demand (device torch ops, the
    same formulas as get_rpe, planners/mind/utils.py:193-212).  The network itself never asks: it evaluates the
    encoding inside its edge-init kernels from the anchors."""
    def __init__(self, ctrs, vecs):
        super().__init__(scene_mask=None)
        self._c, self._v = ctrs, vecs

    def __missing__(self, key):
        if key != "scene":
            raise KeyError(key)
        from .synth import pairwise_rpe
        self["scene"] = pairwise_rpe(self._c, self._v)
        return self["scene"]


class _GeomRPE:
    """data['RPE'] after pre_process when the collated dict carries the anchors the dense encoding was built from
    (TRAJS[b]['TRAJS_CTRS' / 'TRAJS_VECS'], LANE_GRAPH[b]['lane_ctrs' / 'lane_vecs'], scenario_tree.py:176-186):
    only those [sum M_b, 2] arrays are uploaded (~2.6 KB per 160-token scene instead of 512 KB of dense RPE) and
    get_rpe is evaluated on the device.  Behaves like the reference's list of {'scene', 'scene_mask'} dicts; the entries
    are created when somebody indexes or iterates it (the network itself only reads ctrs / vecs / counts)."""

    def __init__(self, ctrs, vecs, counts):
        self.ctrs, self.vecs, self.counts = ctrs, vecs, counts      # [sum M_b, 2] device, per scene [actors ; lanes]
        self._items = None

    def _make(self):
        if self._items is None:
            self._items, off = [], 0
            for n in self.counts:
                self._items.append(_LazyRPE(self.ctrs[off:off + n], self.vecs[off:off + n]))
                off += n
        return self._items

    def __len__(self):
        return len(self.counts)

    def __getitem__(self, i):
        return self._make()[i]

    def __iter__(self):
        return iter(self._make())


class ScenePredNetB200:
    def __init__(self, cfg: Optional[dict], device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ScenePredNetB200 runs on a CUDA device only (no CPU fallback); got %s" % (self.device,))
        if cfg is not None:
            for k, v in _EXPECTED_CFG.items():
                if k in cfg and cfg[k] != v:
                    raise ValueError("unsupported net_cfg[%r]=%r (this build is specialised to %r)" % (k, cfg[k], v))
        self.cfg = cfg
        self._lib = _lib.load()
        h = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(self._lib.mind_create(C.byref(h), idx), "mind_create")
        self._h = h
        self._dev_index = idx
        self._sd: Dict[str, torch.Tensor] = {}
        self._ws = None
        self._out_cache = {}
        self.training = False
        # what the string import gives MIND is the benchmarked tensor-core mode (two tiers, parity <= 1e-3: DESIGN.md 2);
        # net_cfg["precision"] = "fp32" or MIND_B200_PRECISION=fp32 selects the exact SIMT path
        self.set_precision((cfg or {}).get("precision") or os.environ.get("MIND_B200_PRECISION", "f16tc"))
        self.rpe_on_device = True       # pre_process uploads anchors instead of dense RPE when the dict has them
        self._geom_pin = {}

    # ---- nn.Module-like surface used by planners/mind/planner.py:46-49 ----
    def load_state_dict(self, state_dict, strict: bool = True):
        self._sd = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in state_dict.items()}
        for k, v in self._sd.items():
            _lib.check(self._lib.mind_set_weight(self._h, k.encode(), C.c_void_p(v.data_ptr()), v.numel()), k)
        T, Tp = _bezier_bases()
        _lib.check(self._lib.mind_set_weight(self._h, b"__bezier_T", T.ctypes.data_as(C.c_void_p), T.size), "T")
        _lib.check(self._lib.mind_set_weight(self._h, b"__bezier_Tp", Tp.ctypes.data_as(C.c_void_p), Tp.size), "Tp")
        _lib.check(self._lib.mind_finalize_weights(self._h), "mind_finalize_weights")
        return self

    def state_dict(self):
        return dict(self._sd)

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("ScenePredNetB200 cannot be moved off the GPU")
        return self

    def eval(self):
        self.training = False
        return self

    def set_precision(self, mode: str):
        """'fp32' = exact SIMT path; 'f16tc' = tcgen05 fp16-operand path (fp32 accumulate)."""
        self.precision = {"fp32": _lib.PREC_FP32, "f16tc": _lib.PREC_F16TC}[mode]
        _lib.check(self._lib.mind_set_option(self._h, b"precision", self.precision), "precision")
        return self

    def set_option(self, name: str, value: int):
        _lib.check(self._lib.mind_set_option(self._h, name.encode(), int(value)), name)
        return self

    def use_graphs(self, on: bool = True):
        """Capture each (batch shape, pointer set) into a CUDA graph on its second appearance and replay it afterwards."""
        return self.set_option("graph", 1 if on else 0)

    def profile(self, on: bool = True):
        return self.set_option("profile", 1 if on else 0)

    def profile_read(self) -> dict:
        """{tag: (total_ms, count)} of the stage ranges recorded since the last call (synchronises)."""
        buf = C.create_string_buffer(8192)
        _lib.check(self._lib.mind_profile_read(self._h, buf, 8192), "mind_profile_read")
        out = {}
        for line in buf.value.decode().splitlines():
            tag, ms, n = line.split()
            out[tag] = (float(ms), int(n))
        return out

    def sync_check(self):
        _lib.check(self._lib.mind_sync_check(self._h), "mind_sync_check")

    def graph_replays(self) -> int:
        return int(self._lib.mind_graph_replays(self._h))

    def launch_count(self) -> int:
        return int(self._lib.mind_launch_count(self._h))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.mind_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- reference network.py:597-606 ----
    def pre_process(self, data):
        """Host -> device staging of the 7 network inputs.  Same result structure as the reference, different
        mechanics: the per-scene RPE tensors travel in one batched library call into one device buffer
        (mind_upload_packed), and the index lists (only their lengths are used by this network) stay on the host."""
        dev = self.device
        rpe = data["RPE"]
        packed = self._upload_geom(data) if self.rpe_on_device else None
        if packed is None:
            packed = self._upload_rpe(rpe) if isinstance(rpe, (list, tuple)) and len(rpe) > 0 else None
        return (_to_device(data["ACTORS"], dev), data["ACTOR_IDCS"], _to_device(data["LANES"], dev), data["LANE_IDCS"],
                packed if packed is not None else _to_device(rpe, dev),
                _to_device(data["TGT_NODES"], dev), _to_device(data["TGT_RPE"], dev))

    def _upload_geom(self, data):
        """anchors of every scene -> one pinned staging buffer -> one H2D copy; None if the dict does not carry them"""
        trajs, graphs = data.get("TRAJS"), data.get("LANE_GRAPH")
        if not isinstance(trajs, (list, tuple)) or not isinstance(graphs, (list, tuple)) or len(trajs) != len(graphs) or not trajs:
            return None
        try:
            cs, vs, counts = [], [], []
            for t, g in zip(trajs, graphs):
                cs += [t["TRAJS_CTRS"], g["lane_ctrs"]]
                vs += [t["TRAJS_VECS"], g["lane_vecs"]]
                counts.append(int(t["TRAJS_CTRS"].shape[0]) + int(g["lane_ctrs"].shape[0]))
        except (KeyError, TypeError, AttributeError):
            return None
        first = cs[0]
        if not isinstance(first, torch.Tensor) or first.dtype != torch.float32 or first.dim() != 2 or first.shape[1] != 2:
            return None
        m = sum(counts)
        if cs[0].device.type == "cpu":
            slot = self._geom_pin.get(m)
            if slot is None:      # ring of pinned staging buffers: the host may run up to 7 uploads ahead of the copies
                slot = self._geom_pin[m] = [[torch.empty(2, m, 2).pin_memory() for _ in range(8)], 0, [None] * 8]
            k = slot[1] = (slot[1] + 1) % 8
            if slot[2][k] is not None and not slot[2][k].query():
                slot[2][k].synchronize()           # only when the host is 8 uploads ahead of the device
            pin = slot[0][k]
            try:                                  # a tensor of another dtype / shape among the anchors: per-tensor path
                torch.cat(cs, 0, out=pin[0])
                torch.cat(vs, 0, out=pin[1])
            except (RuntimeError, TypeError):
                return None
            devbuf = pin.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            slot[2][k] = ev
        else:
            devbuf = torch.stack([torch.cat(cs, 0), torch.cat(vs, 0)]).to(self.device)
        return _GeomRPE(devbuf[0], devbuf[1], counts)

    def _upload_rpe(self, rpe):
        scenes = []
        for r in rpe:
            t = r["scene"] if isinstance(r, dict) else r
            if (not isinstance(t, torch.Tensor) or t.device.type != "cpu" or t.dtype != torch.float32
                    or t.dim() != 3 or not t.is_contiguous()):
                return None                                   # unusual input: per-tensor path
            scenes.append(t)
        n = len(scenes)
        sizes = (C.c_int64 * n)(*[t.numel() * 4 for t in scenes])
        srcs = (C.c_void_p * n)(*[t.data_ptr() for t in scenes])
        total = self._lib.mind_upload_packed_bytes(sizes, n)
        base = torch.empty(max(int(total), 256), dtype=torch.uint8, device=self.device)
        offs = (C.c_int64 * n)()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mind_upload_packed(srcs, sizes, n, C.c_void_p(base.data_ptr()), base.numel(), offs,
                                                    C.c_void_p(stream)), "mind_upload_packed")
        out = _PackedRPE()
        out.base, out.shapes = base, [tuple(t.shape) for t in scenes]
        bp = base.data_ptr()
        out.ptrs = (C.c_void_p * n)(*[bp + offs[i] for i in range(n)])
        shp0 = out.shapes[0]
        uniform = all(sh == shp0 for sh in out.shapes) and (sizes[0] % 256 == 0)
        if uniform:                                           # one view op for the whole batch
            views = base[:n * sizes[0]].view(torch.float32).view((n,) + shp0).unbind(0)
        else:
            views = [base[offs[i]:offs[i] + sizes[i]].view(torch.float32).view(out.shapes[i]) for i in range(n)]
        for r, v in zip(rpe, views):
            out.append({"scene": v, "scene_mask": r.get("scene_mask") if isinstance(r, dict) else None})
        # pinned sources must outlive the asynchronous copies (torch's caching host allocator does not know about the
        # library's cudaMemcpyAsync): hold them until an event recorded behind the upload has completed
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._pending_host = [(e, src) for e, src in getattr(self, "_pending_host", []) if not e.query()]
        self._pending_host.append((ev, scenes))
        return out

    # ---- reference network.py:582-595 ----
    def forward(self, data):
        packed = self.forward_packed(data)
        self._last_packed = packed
        cls, reg, vel, cov_vel, param, a_off = packed[:6]
        B = cls.shape[0]
        sizes = [a_off[b + 1] - a_off[b] for b in range(B)]  # per-scene views: one split call per output tensor
        res_cls = list(cls.split(1))
        res_reg = list(reg.split(sizes))
        res_aux = [(v, c, p.permute(1, 0, 2, 3)) for v, c, p in zip(vel.split(sizes), cov_vel.split(sizes), param.split(sizes))]
        return res_cls, res_reg, res_aux

    __call__ = forward

    def _workspace(self, nbytes: int):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes * 1.05) + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    def forward_packed(self, data, geom=None, persistent_out=False):
        """Runs the library; returns packed (cls [B,6], reg [A,6,60,5], vel [A,6,60,2],
        cov_vel [A,6,60,3], param [A,6,8,5], actor_offsets list, pack) where `pack` is the one flat buffer cls, reg and
        vel are views of (pack_layout): the send buffer of the path's single all-gather.  `geom` = (ctrs, vecs) device
        tensors [sum M_b, 2] switches RPE evaluation to the device (the 'rpe' entry is ignored).
        `persistent_out` hands out the same output tensors for every call of a given (B, A) shape (the caller must be
        done with the previous result): with option "graph" on, stable pointers let the library replay a captured
        CUDA graph instead of enqueueing ~170 launches."""
        actors, actor_idcs, lanes, lane_idcs, rpe, tgt_nodes, tgt_rpe = data[:7]
        dev = self.device
        B = len(actor_idcs)
        f32 = lambda t: t.to(dev, torch.float32).contiguous()
        actors, lanes, tgt_nodes, tgt_rpe = f32(actors), f32(lanes), f32(tgt_nodes), f32(tgt_rpe)
        a_off = [0] * (B + 1)
        l_off = [0] * (B + 1)
        for b in range(B):
            a_off[b + 1] = a_off[b] + len(actor_idcs[b])
            l_off[b + 1] = l_off[b] + len(lane_idcs[b])
        A, L = a_off[B], l_off[B]
        if actors.shape[0] != A or lanes.shape[0] != L:
            raise ValueError("index lists do not cover ACTORS/LANES contiguously")
        if tgt_nodes.shape[0] != B or tgt_rpe.reshape(B, -1).shape[1] != 20:
            raise ValueError("TGT_NODES / TGT_RPE batch mismatch")
        nmax = max((a_off[b + 1] - a_off[b]) + (l_off[b + 1] - l_off[b]) + 1 for b in range(B))
        bt = _lib.MindBatch()
        bt.n_scenes = B
        ao = (C.c_int32 * (B + 1))(*a_off)
        lo = (C.c_int32 * (B + 1))(*l_off)
        bt.actor_off, bt.lane_off = ao, lo
        bt.actors, bt.lanes = actors.data_ptr(), lanes.data_ptr() if L > 0 else 0
        keep = []
        if geom is None and isinstance(rpe, _GeomRPE) and len(rpe) == B:
            for b in range(B):
                if rpe.counts[b] != (a_off[b + 1] - a_off[b]) + (l_off[b + 1] - l_off[b]):
                    raise ValueError("anchors of scene %d cover %d tokens, expected %d" %
                                     (b, rpe.counts[b], (a_off[b + 1] - a_off[b]) + (l_off[b + 1] - l_off[b])))
            geom = (rpe.ctrs, rpe.vecs)
        if geom is not None:
            ctrs, vecs = f32(geom[0]), f32(geom[1])
            keep += [ctrs, vecs]
            bt.ctrs, bt.vecs = ctrs.data_ptr(), vecs.data_ptr()
            bt.rpe = None
        elif isinstance(rpe, _PackedRPE) and len(rpe) == B and rpe.base.device == dev:
            for b in range(B):
                m = (a_off[b + 1] - a_off[b]) + (l_off[b + 1] - l_off[b])
                if rpe.shapes[b] != (5, m, m):
                    raise ValueError("RPE[%d] has shape %s, expected (5,%d,%d)" % (b, rpe.shapes[b], m, m))
            keep.append(rpe.base)
            bt.rpe = rpe.ptrs
        else:
            ptrs = (C.c_void_p * B)()
            for b in range(B):
                r = rpe[b]["scene"] if isinstance(rpe[b], dict) else rpe[b]
                r = f32(r)
                m = (a_off[b + 1] - a_off[b]) + (l_off[b + 1] - l_off[b])
                if tuple(r.shape) != (5, m, m):
                    raise ValueError("RPE[%d] has shape %s, expected (5,%d,%d)" % (b, tuple(r.shape), m, m))
                keep.append(r)
                ptrs[b] = r.data_ptr()
            bt.rpe = ptrs
        bt.tgt_nodes, bt.tgt_rpe = tgt_nodes.data_ptr(), tgt_rpe.data_ptr()
        outs = self._out_cache.get((B, A)) if persistent_out else None
        if outs is None:
            # cls | reg | vel live in ONE buffer: the decoder kernels write the send buffer of the path's single
            # all-gather directly (mind_b200.distributed.all_gather_packed), no packing copy
            n_cls, n_reg, n_vel = _pad64(B * 6), _pad64(A * 1800), A * 720
            pack = torch.empty(n_cls + n_reg + n_vel, device=dev)
            outs = (pack[:B * 6].view(B, 6), pack[n_cls:n_cls + A * 1800].view(A, 6, 60, 5),
                    pack[n_cls + n_reg:].view(A, 6, 60, 2),
                    torch.empty(A, 6, 60, 3, device=dev), torch.empty(A, 6, 8, 5, device=dev), pack)
            if persistent_out:
                self._out_cache[(B, A)] = outs
        cls, reg, vel, cov_vel, param, self.last_pack = outs
        out = _lib.MindOutputs(cls.data_ptr(), reg.data_ptr(), vel.data_ptr(), cov_vel.data_ptr(), param.data_ptr())
        need = self._lib.mind_workspace_bytes_batch(self._h, C.byref(bt))
        if need < 0:
            _lib.check(1, "mind_workspace_bytes_batch")
        ws = self._workspace(need)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _lib.check(self._lib.mind_forward(self._h, C.byref(bt), C.byref(out), C.c_void_p(ws.data_ptr()),
                                              ws.numel(), C.c_void_p(stream)), "mind_forward")
        self._keep = keep   # inputs must outlive the asynchronous launches
        return cls, reg, vel, cov_vel, param, a_off, self.last_pack

    def debug_tap(self, name: str, numel: int):
        buf = torch.empty(numel, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        n = self._lib.mind_debug_tap(self._h, name.encode(), C.c_void_p(buf.data_ptr()), numel, C.c_void_p(stream))
        if n < 0:
            _lib.check(1, "mind_debug_tap")
        return buf[:n]
