"""Pipelined host <-> device streaming around ``CDSMVSNet``: the way a depth-map job (the loop of the reference's
test.py:197-248, one batch per reference view) is fed on a B200.

One depth map moves ~28 MB of 8-bit images in and ~30 MB of maps out; on one stream that is ~1.2 ms of PCIe time
against ~11 ms of kernels.  Here the upload of work item i+1 and the download of item i-1 run on their own
streams while item i computes, and ``in_flight`` (default 2) work items compute at the same time, each on a cascade
engine and a stream of its own: one map's kernels leave SMs idle (tails of persistent kernels, the small grids of
the deep regulariser levels) that the other map's kernels fill -- measured +6 % maps/s at cfg2 over one map at a time.

    stream = DepthMapStream(model, temperature=0.01)
    pending = []
    for item in work_list:                       # item = (imgs, proj_matrices, depth_values) on the HOST
        pending.append(stream.submit(*item))
        if len(pending) > stream.in_flight:
            consume(stream.result(pending.pop(0)))   # host tensors (pinned), valid until `depth` submits later
    for t in pending:
        consume(stream.result(t))

Every numerical step is still ``CascadeEngine.forward`` (the CUDA kernels); this module only owns streams, events,
staging buffers and the extra engines (which share the model's packed weights).
"""
from __future__ import annotations

import torch


class _Slot:
    def __init__(self):
        self.sig = None           # input signature (shapes, image dtype, projection keys) the buffers below were made for
        self.dev_in = None        # (imgs, proj dict, depth_values) on the device
        self.dev_out = None       # device staging copies of (packed result maps, refined depth | None)
        self.host_in = {}         # persistent pinned staging for unpinned callers
        self.host_pack = None     # pinned copies of the packed result buffer (and the refined depth)
        self.host_out = None      # flat dict of views of host_pack
        self.ev_in = torch.cuda.Event()       # upload finished
        self.ev_done = torch.cuda.Event()     # compute + staging copy finished (inputs may be overwritten)
        self.ev_out = torch.cuda.Event()      # download finished
        self.used = False


def _flatten(out):
    flat = {}
    for k, v in out.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                flat[f"{k}.{kk}"] = vv
    # refine=True: the full-resolution refined depth is a map of its own (refine=False aliases the last stage's depth)
    rd = out.get("refined_depth")
    if rd is not None and all(rd.data_ptr() != v.data_ptr() for v in flat.values()):
        flat["refined_depth"] = rd
    return flat


class DepthMapStream:
    def __init__(self, model, temperature=0.001, depth=None, use_graph=True, in_flight=2):
        self.model = model
        self.use_graph = use_graph      # replay the forward as one CUDA graph (captured on the first item of a shape)
        self.temperature = float(temperature)
        self.in_flight = max(1, int(in_flight))            # work items computing at the same time (one engine + stream each)
        depth = self.in_flight + 1 if depth is None else int(depth)
        if depth < self.in_flight:
            raise ValueError("depth (buffered work items) must be at least in_flight")
        self.slots = [_Slot() for _ in range(depth)]
        self.n = 0
        self.s_in = torch.cuda.Stream()
        self.s_out = torch.cuda.Stream()
        self.s_comp = [torch.cuda.Stream() for _ in range(self.in_flight)] if self.in_flight > 1 else [None]
        self._base = None               # the model's own engine the extra ones were cloned from
        self._extra = []

    def _engine(self, k, dev):
        """Engine of compute lane k: the model's own for lane 0, clones over the same packed weights for the others (re-made
        when the model re-packs its weights)."""
        base = self.model.engine(dev)
        if base is not self._base:
            self._base, self._extra = base, [base.clone() for _ in range(self.in_flight - 1)]
        return base if k == 0 else self._extra[k - 1]

    @staticmethod
    def _signature(imgs, proj_matrices, depth_values):
        return (tuple(imgs.shape), imgs.dtype, tuple(sorted((k, tuple(v.shape)) for k, v in proj_matrices.items())),
                tuple(depth_values.shape))

    def _stage_host(self, slot, name, t, dtype):
        """A pinned host tensor holding ``t``: ``t`` itself if the caller pinned it, else the slot's persistent pinned staging
        buffer (allocated once per input signature -- cudaHostAlloc per submit would serialise the pipeline)."""
        t = t if t.dtype == dtype else t.to(dtype)
        if t.is_pinned():
            return t
        buf = slot.host_in.get(name)
        if buf is None or buf.shape != t.shape or buf.dtype != dtype:
            buf = slot.host_in[name] = torch.empty(t.shape, dtype=dtype).pin_memory()
        buf.copy_(t)
        return buf

    def submit(self, imgs, proj_matrices, depth_values):
        """Enqueue one work item given as HOST tensors; returns a ticket for ``result``.  ``imgs`` may be fp32 in [0,1] or
        uint8 (the bytes the data layer read; divided by 255 on the device, a quarter of the upload)."""
        if imgs.is_cuda:
            raise ValueError("DepthMapStream.submit takes host tensors (use model(...) for device-resident inputs)")
        dev = next(self.model.parameters()).device
        lane = self.n % self.in_flight
        caller = torch.cuda.current_stream(dev)
        compute = caller if self.in_flight == 1 else self.s_comp[lane]
        if compute is not caller:
            compute.wait_stream(caller)     # work the caller queued before this submit stays ahead of it
        slot = self.slots[self.n % len(self.slots)]
        img_dtype = torch.uint8 if imgs.dtype == torch.uint8 else torch.float32
        sig = self._signature(imgs, proj_matrices, depth_values)
        if slot.sig != sig:
            # any change of shape, image dtype or projection keys re-creates the slot's device inputs (a stale key from a
            # previous signature must never reach the engine) and its staging buffers
            if slot.used:
                slot.ev_out.synchronize()
            slot.sig, slot.host_in, slot.dev_out, slot.host_out = sig, {}, None, None
            slot.dev_in = (torch.empty(imgs.shape, dtype=img_dtype, device=dev),
                           {k: torch.empty(v.shape, dtype=torch.float32, device=dev) for k, v in proj_matrices.items()},
                           torch.empty(depth_values.shape, dtype=torch.float32, device=dev))
        if slot.used:
            slot.ev_in.synchronize()   # the previous upload out of this slot's pinned staging has finished
        imgs = self._stage_host(slot, "imgs", imgs, img_dtype)
        depth_values = self._stage_host(slot, "dv", depth_values, torch.float32)
        proj_matrices = {k: self._stage_host(slot, "proj." + k, v, torch.float32) for k, v in proj_matrices.items()}
        # ---- upload on its own stream, once the slot's previous occupant has been consumed by the kernels
        with torch.cuda.stream(self.s_in):
            if slot.used:
                self.s_in.wait_event(slot.ev_done)
            slot.dev_in[0].copy_(imgs, non_blocking=True)
            for k, v in proj_matrices.items():
                slot.dev_in[1][k].copy_(v, non_blocking=True)
            slot.dev_in[2].copy_(depth_values, non_blocking=True)
            slot.ev_in.record(self.s_in)
        # ---- kernels: on the caller's stream (in_flight = 1) or on this lane's stream and engine
        compute.wait_event(slot.ev_in)
        engine = self._engine(lane, dev)
        run = engine.forward_graph if self.use_graph else engine.forward
        with torch.cuda.stream(compute):
            out = run(slot.dev_in[0], slot.dev_in[1], slot.dev_in[2], self.temperature)
            self._stage_out(slot, engine, out, compute)
        # ---- download on its own stream
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot.ev_done)
            for d, h in zip(slot.dev_out, slot.host_pack):
                if d is not None:
                    h.copy_(d, non_blocking=True)
            slot.ev_out.record(self.s_out)
        slot.used = True
        ticket = self.n
        self.n += 1
        return ticket

    def _stage_out(self, slot, engine, out, compute):
        # the engine keeps every result map in ONE packed buffer (+ the refined depth when refine=True): one staging copy and
        # one download per item; the host-side dict is a set of views of the pinned copy
        pack = engine._out_pack
        refined = out["refined_depth"] if getattr(self.model, "refine", False) else None
        if slot.dev_out is None:
            slot.dev_out = (torch.empty_like(pack), torch.empty_like(refined) if refined is not None else None)
            slot.host_pack = (torch.empty(pack.shape, dtype=pack.dtype).pin_memory(),
                              torch.empty(refined.shape, dtype=refined.dtype).pin_memory() if refined is not None else None)
            slot.host_out = _flatten(engine.outputs_from(*slot.host_pack))
        if slot.used:
            compute.wait_event(slot.ev_out)      # the staging tensors were last read by the slot's previous download
        slot.dev_out[0].copy_(pack, non_blocking=True)   # engine buffers are reused by the next forward: stage the results
        if refined is not None:
            slot.dev_out[1].copy_(refined, non_blocking=True)
        slot.ev_done.record(compute)

    def result(self, ticket):
        """Block until work item ``ticket`` is on the host; returns {"stageK.depth" | ".photometric_confidence" | ".norm_curv"}
        pinned host tensors, valid until ``depth`` (default in_flight + 1) further submits."""
        if ticket < self.n - len(self.slots) or ticket >= self.n:
            raise ValueError(f"ticket {ticket} is no longer (or not yet) buffered")
        slot = self.slots[ticket % len(self.slots)]
        slot.ev_out.synchronize()
        return slot.host_out

    def bytes_per_item(self):
        slot = next(s for s in self.slots if s.used)
        h2d = (slot.dev_in[0].numel() * slot.dev_in[0].element_size() + slot.dev_in[2].numel() * 4 +
               sum(v.numel() * 4 for v in slot.dev_in[1].values()))
        d2h = sum(v.numel() * v.element_size() for v in slot.host_out.values())
        return h2d, d2h
