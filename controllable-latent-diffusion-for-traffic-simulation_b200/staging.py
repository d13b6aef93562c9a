"""Host -> device staging of per-call inputs for the end-to-end path (PyTorch plumbing only: streams, events, pinned memory).

`HostStager` double-buffers the input dictionaries of consecutive `DmModel.forward` calls: while call i runs on the caller's
stream, the pinned host tensors of call i + 1 are copied into the other device buffer set on a dedicated copy stream, so the
PCIe transfer (drivable maps, neighbour futures, cond features: tens to hundreds of MB per call) overlaps the sampler instead of
preceding it.  Results are read back with `read_back` into pinned host buffers on the compute stream.

    st = HostStager(device)
    st.put(batch_host, aux_host)                  # first call's inputs
    for i in range(n):
        batch_d, aux_d, slot = st.get()           # compute stream waits for copy i
        if i + 1 < n: st.put(batch_host, aux_host)   # copy i + 1 overlaps call i
        out = dm(batch_d, aux_d, algo, ...)
        st.release(slot)                          # the buffer set may be overwritten once call i has finished
        host = st.read_back(out, ("traj", "offroad", "coll"))
    st.finish()
"""
import torch


class HostStager:
    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.depth = int(depth)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots = [dict(bufs=None, ready=torch.cuda.Event(), free=torch.cuda.Event(), used=False) for _ in range(self.depth)]
        self._put, self._got = 0, 0
        self._out = {}
        self._out_evt = None

    @staticmethod
    def _dev_like(d, device):
        return {k: (torch.empty(v.shape, dtype=v.dtype, device=device) if torch.is_tensor(v) else v) for k, v in d.items()}

    def put(self, *host_dicts):
        """Enqueue the copy of one call's inputs (dicts of pinned host tensors) on the copy stream."""
        slot = self._slots[self._put % self.depth]
        self._put += 1
        if slot["bufs"] is None:
            slot["bufs"] = [self._dev_like(d, self.device) for d in host_dicts]
        if slot["used"]:
            self.copy_stream.wait_event(slot["free"])          # the call that read this buffer set has finished
        with torch.cuda.stream(self.copy_stream):
            for dst, src in zip(slot["bufs"], host_dicts):
                for k, v in src.items():
                    if torch.is_tensor(v):
                        dst[k].copy_(v, non_blocking=True)
            slot["ready"].record(self.copy_stream)

    def get(self):
        """-> (*device dicts, slot id): the compute stream waits until the copy has landed."""
        idx = self._got % self.depth
        slot = self._slots[idx]
        self._got += 1
        torch.cuda.current_stream(self.device).wait_event(slot["ready"])
        return (*slot["bufs"], idx)

    def release(self, idx):
        slot = self._slots[idx]
        slot["free"].record(torch.cuda.current_stream(self.device))
        slot["used"] = True

    def read_back(self, out, keys):
        """Device -> pinned host copy of the named results on the compute stream; waits for the PREVIOUS call's copy first, so
        every call's results are complete in host memory one call later (and all of them after `finish`)."""
        if self._out_evt is not None:
            self._out_evt.synchronize()
        host = {}
        for k in keys:
            v = out[k]
            buf = self._out.get(k)
            if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                buf = self._out[k] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
            buf.copy_(v, non_blocking=True)
            host[k] = buf
        self._out_evt = torch.cuda.Event()
        self._out_evt.record(torch.cuda.current_stream(self.device))
        return host

    def finish(self):
        if self._out_evt is not None:
            self._out_evt.synchronize()
        self.copy_stream.synchronize()
