"""Scan pipeline: CUDA-graph replays of the hot path with several (default four) scans in flight per stream.

Dependencies between consecutive scans of ONE stream are exactly the reference's
(networks/multi_view_encoder.py:433-439, voxel_voting.py:140,182):
  * the cascade projection of a scan depends on nothing from earlier scans;
  * temporal fusion of scan t reads the short-term memory written by scan t-1;
  * long-term voting of scan t follows the network of scan t and the voting of scan t-1 (ring buffer).
So three graphs are captured per scan — P (5 pools + 5 gathers), M (2 deformable-attention layers + memory
update), V (voxel + instance voting) — and replayed on three CUDA streams: even scans' P/M on stream A, odd
scans' on stream B, ... (one P/M stream per scan in flight), all V on one more stream, ordered with events. While scan
t sits in its latency-bound small kernels, the HBM-bound pooling of the following scans keeps the memory system busy
(measured on B200, scans/s device-resident / end to end: 1 in flight 3424 / 2741, 2: 4404 / 3137, 4: 4680 / 3355,
8: 4535 / 3230). Results are identical to the
serial step (tests/test_gpu_parity.py::test_pipeline_matches_serial_step).
"""
import torch


class ScanPipeline:
    def __init__(self, hot, dev_scans, use_graphs=True, scans_in_flight=4):
        assert len(dev_scans) % scans_in_flight == 0, "resident scan buffers must be a multiple of scans_in_flight"
        if use_graphs:  # a captured voting graph bakes its ring slot (scan index mod 8) in
            from .stream import HISTORY
            assert len(dev_scans) % HISTORY == 0, "graph mode needs a multiple of %d scan buffers" % HISTORY
        self.hot, self.scans, self.n = hot, dev_scans, len(dev_scans)
        dev = hot.device
        import os
        # the voting stream runs at high priority: same device-resident throughput, +1.5 % end to end (the labels
        # leave for the host sooner); SMOS_PIPE_PRIO=equal / v_low select the other arrangements that were measured
        prio = os.environ.get("SMOS_PIPE_PRIO", "v_high")
        p_pm, p_v = (0, -1) if prio == "v_high" else ((-1, 0) if prio == "v_low" else (0, 0))
        self.pm_streams = [torch.cuda.Stream(dev, priority=p_pm) for _ in range(scans_in_flight)]
        self.sC = torch.cuda.Stream(dev, priority=p_v)
        self.use_graphs = use_graphs
        self.gP, self.gM, self.gV = [None] * self.n, [None] * self.n, [None] * self.n
        self.proj, self.out = [None] * self.n, [None] * self.n
        self.m_done = [torch.cuda.Event() for _ in range(self.n)]
        self.v_done = [torch.cuda.Event() for _ in range(self.n)]
        self.submitted = 0
        hot.overlap_voting = False
        # several scans in flight: the latency-bound PointNet stem (raw-scan batches) leaves ~40 % of the SMs to the
        # HBM-bound kernels of the neighbouring scans (measured on B200, end to end: 148 CTAs 3426-3446 scans/s,
        # 84-96 CTAs 3484-3597); SMOS_PIPE_STEM_SHARE overrides
        if scans_in_flight > 1 and getattr(hot, "stem_sm_share", None) == 1.0:
            hot.stem_sm_share = float(os.environ.get("SMOS_PIPE_STEM_SHARE", "0.61"))
        with torch.no_grad():
            # eager warm-up of every buffer / ring slot (also allocates persistent scratch outside capture)
            for j in range(self.n):
                hot.scan_index = j
                with torch.cuda.stream(self._pm_stream(j)):
                    self.proj[j] = hot.projection(dev_scans[j])
                    hot.temporal_fusion(dev_scans[j])
                    self.out[j] = hot.long_term_voting(dev_scans[j])
                torch.cuda.synchronize(dev)
            if use_graphs:
                pools = {id(st): torch.cuda.graph_pool_handle() for st in self.pm_streams}
                pool_c = torch.cuda.graph_pool_handle()
                for j in range(self.n):
                    s = self._pm_stream(j)
                    self.gP[j] = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.gP[j], pool=pools[id(s)], stream=s):
                        self.proj[j] = hot.projection(dev_scans[j])
                    self.gM[j] = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.gM[j], pool=pools[id(s)], stream=s):
                        hot.temporal_fusion(dev_scans[j])
                    hot.scan_index = j
                    self.gV[j] = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.gV[j], pool=pool_c, stream=self.sC):
                        self.out[j] = hot.long_term_voting(dev_scans[j])
                torch.cuda.synchronize(dev)
        hot.scan_index = 0

    def _pm_stream(self, j):
        return self.pm_streams[j % len(self.pm_streams)]

    def streams(self):
        return tuple(self.pm_streams) + (self.sC,)

    def submit(self, ready_event=None, also_ready=()):
        """Enqueue the next scan (buffer index = scan number mod n). `ready_event`: its inputs are resident;
        `also_ready`: events of other buffers this scan reads (the older raw scans of a resident window)."""
        i = self.submitted
        j = i % self.n
        s = self._pm_stream(j)
        hot = self.hot
        with torch.no_grad(), torch.cuda.stream(s):
            if ready_event is not None:
                s.wait_event(ready_event)
            for e in also_ready:
                s.wait_event(e)
            if self.use_graphs:
                self.gP[j].replay()
            else:
                self.proj[j] = hot.projection(self.scans[j])
            if i > 0:
                s.wait_event(self.m_done[(i - 1) % self.n])   # short-term memory written by the previous scan
            if self.use_graphs:
                self.gM[j].replay()
            else:
                hot.temporal_fusion(self.scans[j])
            self.m_done[j].record(s)
        with torch.no_grad(), torch.cuda.stream(self.sC):
            if ready_event is not None:
                self.sC.wait_event(ready_event)
            self.sC.wait_event(self.m_done[j])                # voting post-processes this scan's predictions
            if self.use_graphs:
                self.gV[j].replay()
            else:
                hot.scan_index = i
                self.out[j] = hot.long_term_voting(self.scans[j])
            self.v_done[j].record(self.sC)
        self.submitted += 1
        hot.scan_index = self.submitted
        return j

    def join(self, stream):
        """Make `stream` wait for everything submitted so far."""
        for e in self.m_done + self.v_done:
            stream.wait_event(e)
