"""Host-buffer feature extraction: the call a user with data in host memory makes.

``FeaturePipeline.run(x_host)`` takes [n_series, len] host data -- float32 samples, or the int16 raw counts a
wearable logger stores (``count_scale`` turns a count into a sample; 2 bytes per sample cross PCIe and the widening
happens on the device, exactly) -- as numpy or a pinned torch tensor, streams it through the GPU in chunks of whole series -- H2D copy of chunk i+1 and D2H copy
of table i-1 overlap the kernels of chunk i (two copy streams + one compute stream, double-buffered
device and pinned staging buffers) -- and returns the [n_series, nw, n_features] table on the host.
torch provides the buffers / streams / events; every number comes from libmhb200's kernels.
"""
import numpy as np

from . import engine


class FeaturePipeline:
    def __init__(self, features, wsize, wstep, fs=1.0, zc_threshold=0.0, chunk_series=12, out_float32=True,
                 count_scale=1.0):
        self.features = list(features)
        self.count_scale = float(count_scale)
        self.wsize, self.wstep, self.fs, self.zc = int(wsize), int(wstep), float(fs), float(zc_threshold)
        self.chunk_series = int(chunk_series)
        self.out_float32 = out_float32
        self._bufs = None

    def _ensure(self, torch, n, nw, device, raw):
        key = (n, nw, str(device), raw)
        if self._bufs is not None and self._bufs["key"] == key:
            return self._bufs
        cs, nf = self.chunk_series, len(self.features)
        odt = torch.float32 if self.out_float32 else torch.float64
        b = {"key": key,
             "din": [torch.empty((cs, n), dtype=torch.float32, device=device) for _ in range(2)],
             "draw": [torch.empty((cs, n), dtype=torch.int16, device=device) for _ in range(2)] if raw else None,
             "dout": [torch.empty((cs, nw, nf), dtype=odt, device=device) for _ in range(2)],
             "h2d": torch.cuda.Stream(device), "d2h": torch.cuda.Stream(device), "comp": torch.cuda.Stream(device),
             "in_free": [torch.cuda.Event() for _ in range(2)], "in_ready": [torch.cuda.Event() for _ in range(2)],
             "out_ready": [torch.cuda.Event() for _ in range(2)], "out_free": [torch.cuda.Event() for _ in range(2)]}
        self._bufs = b
        return b

    def run(self, x_host, out_host=None, device=None):
        torch = engine.require_cuda()
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        if isinstance(x_host, torch.Tensor):
            xt = x_host
        else:
            a = np.asarray(x_host)
            xt = torch.from_numpy(np.ascontiguousarray(a if a.dtype == np.int16 else a.astype(np.float32, copy=False)))
        if xt.dim() != 2 or xt.dtype not in (torch.float32, torch.int16) or xt.is_cuda:
            raise ValueError("FeaturePipeline.run takes a host float32 or int16 [n_series, len] array")
        raw = xt.dtype == torch.int16
        ns, n = xt.shape
        nw = engine.n_windows(n, self.wsize, self.wstep)
        nf = len(self.features)
        odt = torch.float32 if self.out_float32 else torch.float64
        if out_host is None:
            out_host = torch.empty((ns, nw, nf), dtype=odt).pin_memory()
        b = self._ensure(torch, n, nw, device, raw)
        lib = engine.L.load()
        cs = self.chunk_series
        nchunks = (ns + cs - 1) // cs
        cur = torch.cuda.current_stream(device)
        for s in (b["h2d"], b["comp"], b["d2h"]):
            s.wait_stream(cur)
        for i in range(nchunks):
            k = i & 1
            a, e = i * cs, min(ns, (i + 1) * cs)
            m = e - a
            with torch.cuda.stream(b["h2d"]):
                if i >= 2:
                    b["h2d"].wait_event(b["in_free"][k])          # kernels of chunk i-2 are done with this buffer
                (b["draw"] if raw else b["din"])[k][:m].copy_(xt[a:e], non_blocking=True)
                b["in_ready"][k].record(b["h2d"])
            with torch.cuda.stream(b["comp"]):
                b["comp"].wait_event(b["in_ready"][k])
                if i >= 2:
                    b["comp"].wait_event(b["out_free"][k])        # table of chunk i-2 has left the device
                if raw:     # counts -> float32 samples (exact), csrc/series_reduce.cu
                    engine.L.check(lib.mhb_widen_i16_f32(b["draw"][k].data_ptr(), m * n, self.count_scale, b["din"][k].data_ptr(),
                                                         engine._stream_ptr(torch)), "widen_i16")
                engine.window_table(b["din"][k][:m], self.wsize, self.wstep, self.features, zc_threshold=self.zc,
                                    fs=self.fs, out=b["dout"][k][:m])
                b["in_free"][k].record(b["comp"])
                b["out_ready"][k].record(b["comp"])
            with torch.cuda.stream(b["d2h"]):
                b["d2h"].wait_event(b["out_ready"][k])
                out_host[a:e].copy_(b["dout"][k][:m], non_blocking=True)
                b["out_free"][k].record(b["d2h"])
        cur.wait_stream(b["d2h"])
        cur.wait_stream(b["comp"])
        # the table is HOST memory: it is valid for the caller only once the last device-to-host copy has landed, so
        # the host waits here (stream-order alone would let `out_host.numpy()` read rows that are still in flight)
        b["d2h"].synchronize()
        return out_host

    def launches_per_chunk(self):
        """Kernels of this library per chunk: one per feature family present (+ the widening kernel for raw counts)."""
        return len({f.family for f in self.features})
