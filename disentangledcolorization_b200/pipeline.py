"""Double-buffered host<->device pipeline around `AnchorColorProb.forward` (the host I/O
around the path -- reference main/colorizer/inference.py:86-139 moves one image per iteration and synchronises on
every `.cpu()`).

Step i's inputs are copied from pinned host memory on a copy stream while step i-1 computes, and step i's result
travels back on a second copy stream while step i+1 computes.  Every step still pays its own H2D and D2H transfers;
they overlap the forward instead of serialising with it (PCIe is full duplex and the forward is ~14 ms against
~1.7 ms of copies at batch 64, 256x256).

Batches may differ in shape (the CLI's --no_resize mode): device and pinned buffers are kept per slot and re-allocated
when a slot meets a new shape.  `post(out_tuple, gray, ab) -> tensor` (optional) runs on the compute stream right after
the forward and names what travels back instead of pred_colors (the CLI converts Lab -> RGB uint8 on the device there).
"""
import torch


class ColorizePipeline:
    def __init__(self, model, batch=None, height=None, width=None, device=None, depth=2, sampled_T=0):
        self.model = model
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.depth = depth
        self.sampled_T = sampled_T
        self.forward_kwargs = None       # optional callable -> extra keyword arguments of each forward (e.g. init_idx of a shard)
        self.h2d = torch.cuda.Stream(self.dev)
        self.d2h = torch.cuda.Stream(self.dev)
        self.slots = [dict(shape=None, gray=None, ab=None, out_host=None, ev_in=torch.cuda.Event(), ev_compute=torch.cuda.Event(),
                           ev_out=torch.cuda.Event(), used=False) for _ in range(depth)]
        self.h2d_bytes = self.d2h_bytes = 0
        if batch is not None:            # fixed-shape use (bench.py): allocate up front
            for s in self.slots:
                self._fit(s, (batch, height, width))
                s["out_host"] = torch.empty(batch, 2, height, width).pin_memory()
            self.h2d_bytes = batch * 3 * height * width * 4
            self.d2h_bytes = batch * 2 * height * width * 4

    def _fit(self, s, shape):
        if s["shape"] != shape:
            if s["used"]:                 # the slot's last forward and D2H are done before its buffers are replaced
                s["ev_compute"].synchronize()
                s["ev_out"].synchronize()
            b, h, w = shape
            s["gray"] = torch.empty(b, 1, h, w, device=self.dev)
            s["ab"] = torch.empty(b, 2, h, w, device=self.dev)
            s["shape"] = shape

    def _stage(self, slot, gray_host, ab_host, host_free=None):
        """`host_free` (optional torch.cuda.Event): recorded once both copies have read the pinned source buffers."""
        s = self.slots[slot]
        self._fit(s, (gray_host.shape[0], gray_host.shape[2], gray_host.shape[3]))
        if s["used"]:
            self.h2d.wait_event(s["ev_compute"])      # the forward that last read this slot's inputs has finished
        with torch.cuda.stream(self.h2d):
            s["gray"].copy_(gray_host, non_blocking=True)
            s["ab"].copy_(ab_host, non_blocking=True)
            s["ev_in"].record(self.h2d)
            if host_free is not None:
                host_free.record(self.h2d)

    def run(self, batches, on_step=None, before_step=None, on_result=None, keep="copy", post=None):
        """batches: iterable of (gray_host, ab_host[, host_free_event]) pinned fp32 tensors (any length, shapes may change).
        `on_step(out_tuple)` runs on the compute stream right after each forward (e.g. the all-gather of the multi-GPU
        job); `post(out_tuple, gray_dev, ab_dev)` returns the device tensor to copy back (default: pred_colors).

        Results: `on_result(i, host_tensor)` (when given) is called for every step as soon as its D2H copy has landed,
        while the slot buffer is still that step's.  The returned list holds one pinned host tensor per step:
        keep="copy" (default) hands out private copies; keep="alias" returns the `depth` slot buffers themselves, each
        valid only until `depth` steps later; keep="none" returns nothing (results consumed through `on_result`)."""
        if keep not in ("copy", "alias", "none"):
            raise ValueError("keep must be 'copy', 'alias' or 'none'")
        main = torch.cuda.current_stream(self.dev)
        results = []
        static_out = bool(getattr(self.model, "use_cuda_graph", False) and getattr(self.model, "graph_static_outputs", False))
        last_out_ev = None

        def retire(j):
            """Step j's D2H is complete (waits for it): deliver / copy its result before the slot is reused."""
            sj = self.slots[j % self.depth]
            sj["ev_out"].synchronize()
            if on_result is not None:
                on_result(j, sj["out_host"])
            if keep != "none":
                results.append(sj["out_host"].clone() if keep == "copy" else sj["out_host"])

        it = iter(batches)
        nxt = next(it, None)
        if nxt is None:
            return results
        self._stage(0, *nxt[:3])
        i = 0
        while nxt is not None:
            s = self.slots[i % self.depth]
            nxt = next(it, None)
            if nxt is not None:
                self._stage((i + 1) % self.depth, *nxt[:3])
            if i >= self.depth:
                retire(i - self.depth)                  # this slot's previous result leaves before its host buffer is reused
                                                        # (that step finished while step i-1 was running: no bubble)
            main.wait_event(s["ev_in"])
            if static_out and last_out_ev is not None:
                # graph mode with static outputs: pred_colors lives in ONE graph-owned buffer that the next replay
                # overwrites, so the forward must not start before the previous step's D2H has read it
                main.wait_event(last_out_ev)
            if before_step is not None:
                before_step()
            kw = self.forward_kwargs() if self.forward_kwargs is not None else {}
            out = self.model(s["gray"], s["ab"], True, self.sampled_T, **kw)
            if on_step is not None:
                on_step(out)
            back = out[2] if post is None else post(out, s["gray"], s["ab"])
            s["ev_compute"].record(main)
            s["used"] = True
            if s["out_host"] is None or s["out_host"].shape != back.shape or s["out_host"].dtype != back.dtype:
                s["out_host"] = torch.empty(back.shape, dtype=back.dtype).pin_memory()
            self.d2h.wait_event(s["ev_compute"])
            with torch.cuda.stream(self.d2h):
                s["out_host"].copy_(back, non_blocking=True)
                s["ev_out"].record(self.d2h)
            last_out_ev = s["ev_out"]
            if not (static_out and post is None):
                back.record_stream(self.d2h)
            i += 1
        n = i
        for j in range(max(0, n - self.depth), n):      # retire() has run for steps [0, n - depth)
            retire(j)
        self.d2h.synchronize()
        main.synchronize()
        return results
