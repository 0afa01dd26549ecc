"""Double-buffered host<->device pipeline around `AnchorColorProb.forward` (the host I/O
around the path -- reference main/colorizer/inference.py:86-139 moves one image per iteration and synchronises on
every `.cpu()`).

Step i's inputs are copied from pinned host memory on a copy stream while step i-1 computes, and step i's
`pred_colors` travels back on a second copy stream while step i+1 computes.  Every step still pays its own H2D and
D2H transfers; they overlap the forward instead of serialising with it (PCIe is full duplex and the forward is
~16 ms against ~1.7 ms of copies at batch 64, 256x256)."""
import torch


class ColorizePipeline:
    def __init__(self, model, batch, height, width, device=None, depth=2, sampled_T=0):
        self.model = model
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.depth = depth
        self.sampled_T = sampled_T
        self.h2d = torch.cuda.Stream(self.dev)
        self.d2h = torch.cuda.Stream(self.dev)
        self.slots = []
        for _ in range(depth):
            self.slots.append(dict(
                gray=torch.empty(batch, 1, height, width, device=self.dev),
                ab=torch.empty(batch, 2, height, width, device=self.dev),
                out_host=torch.empty(batch, 2, height, width).pin_memory(),
                ev_in=torch.cuda.Event(), ev_compute=torch.cuda.Event(), ev_out=torch.cuda.Event(), used=False))
        self.h2d_bytes = batch * 3 * height * width * 4
        self.d2h_bytes = batch * 2 * height * width * 4

    def _stage(self, slot, gray_host, ab_host):
        s = self.slots[slot]
        if s["used"]:
            self.h2d.wait_event(s["ev_compute"])      # the forward that last read this slot's inputs has finished
        with torch.cuda.stream(self.h2d):
            s["gray"].copy_(gray_host, non_blocking=True)
            s["ab"].copy_(ab_host, non_blocking=True)
            s["ev_in"].record(self.h2d)

    def run(self, batches, on_step=None, before_step=None):
        """batches: sequence of (gray_host, ab_host) pinned fp32 tensors.  Returns the list of pinned host tensors that
        received pred_colors of each step (slot buffers: valid until `depth` steps later).  `on_step(out_tuple)` runs on
        the compute stream right after each forward (e.g. the all-gather of the multi-GPU job)."""
        main = torch.cuda.current_stream(self.dev)
        n = len(batches)
        results = []
        if n == 0:
            return results
        self._stage(0, *batches[0])
        for i in range(n):
            s = self.slots[i % self.depth]
            if i + 1 < n:
                self._stage((i + 1) % self.depth, *batches[i + 1])
            main.wait_event(s["ev_in"])
            if s["used"]:
                main.wait_event(s["ev_out"])            # (no-op in practice) previous D2H out of this slot is done
            if before_step is not None:
                before_step()
            out = self.model(s["gray"], s["ab"], True, self.sampled_T)
            if on_step is not None:
                on_step(out)
            s["ev_compute"].record(main)
            s["used"] = True
            self.d2h.wait_event(s["ev_compute"])
            with torch.cuda.stream(self.d2h):
                s["out_host"].copy_(out[2], non_blocking=True)
                s["ev_out"].record(self.d2h)
            out[2].record_stream(self.d2h)
            results.append(s["out_host"])
        self.d2h.synchronize()
        main.synchronize()
        return results
