"""Frame driver: the reference's evaluation loop (`eval_seq`, test_LD_37.py:118-181) as a device-resident pipeline.

What the reference does per output frame i of a sequence with T frames (all on the host, one frame at a time):
  o_list = clip(i-3 .. i+3, 0, T-1)                                   generate_input_index   test_LD_37.py:13-16
  LR / partition-map planes: uint8 / 255, 270-row frames + 2 zero rows generate_input / _PM    :19-46  (side info of frame max(1, j))
  residual / unfiltered planes / 255                                    generate_RM / _UF       :49-74
  mvs = mv2mvs(mvl0[max(1, i)]); modify_mv_for_end_frames(i, mvs, T)    :83-105, :157-163, :209-234
  sr, L1_fea = model(..., L1_fea)  (cached features after frame 0)      :165-169
  crop the padded rows, clamp * 255 -> uint8 -> PNG                     :172-180
  PSNR / SSIM of the PNGs against the ground truth, averaged per frame  metric/psnr_ssim.py:446-484

Here: S sequences of equal shape are batched; only the NEW frame of each window crosses PCIe (uint8 planes + the int MV
field from double-buffered pinned staging, on a copy stream that runs one step ahead of the compute stream); the
7-frame windows, the decoded flows, the feature cache and the SR frame stay in HBM; the uint8 SR frame returns through
pinned memory one step later; PSNR / SSIM are computed on the GPU (csrc/metrics.cu) into a per-sequence [n, 3] fp64
accumulator that `sharding.gather_metrics` reduces over ranks at the end of the job.  The Gumbel noise of
LLongRangAttention is keyed by (seed, sequence id, frame, neighbour), so results do not depend on batching or sharding.
"""
import os

import numpy as np
import torch

from . import metrics, priors, sharding
from .graph import GraphedStep

NEIGHBOURS = (0, 1, 2, 4, 5, 6)


def generate_input_index(center_index, frame_number, max_index):
    """Window of `frame_number` frame indices centred on `center_index`, clipped to [0, max_index] (test_LD_37.py:13-16)."""
    lo = center_index - frame_number // 2
    return [min(max(j, 0), max_index) for j in range(lo, lo + frame_number)]


def side_info_index(j):
    """The reference reads the side information of frame max(1, j): frame 0 has none (test_LD_37.py:36,54,67,155)."""
    return max(1, j)


def new_frame_of_step(i, n_frames):
    """Index of the one frame that enters the window at step i >= 1 (= last entry of generate_input_index)."""
    return min(i + 3, n_frames - 1)


class Sequence:
    """One coded sequence as the reference stores it on disk, held as integer arrays.

    lr, pm, unflt: uint8 [T, h, W];  res: integer [T, h, W] (first channel of *_res.npy);  mvl0: int8 / int32 [T, h, W, 3]
    (mv_a, mv_b, ref-distance);  gt: uint8 [T, 4h, 4W] or None.  Entry 0 of the side-information arrays is never read."""

    def __init__(self, lr, pm, res, unflt, mvl0, gt=None, name="seq"):
        self.lr = np.ascontiguousarray(lr, dtype=np.uint8)
        self.pm = np.ascontiguousarray(pm, dtype=np.uint8)
        self.unflt = np.ascontiguousarray(unflt, dtype=np.uint8)
        res = np.asarray(res)
        if not np.issubdtype(res.dtype, np.integer) or res.min(initial=0) < -32768 or res.max(initial=0) > 32767:
            raise ValueError("Sequence: residual maps must be integers within int16")
        self.res = np.ascontiguousarray(res, dtype=np.int16)
        mv = np.asarray(mvl0)
        if not np.issubdtype(mv.dtype, np.integer):
            raise ValueError("Sequence: mvl0 must be an integer array [T, h, W, 3]")
        small = mv.min(initial=0) >= -128 and mv.max(initial=0) <= 127
        self.mvl0 = np.ascontiguousarray(mv, dtype=np.int8 if small else np.int32)
        self.gt = None if gt is None else np.ascontiguousarray(gt, dtype=np.uint8)
        self.name = name
        T, h, W = self.lr.shape
        for a, what in ((self.pm, "pm"), (self.res, "res"), (self.unflt, "unflt")):
            if a.shape != (T, h, W):
                raise ValueError("Sequence: %s has shape %s, expected %s" % (what, a.shape, (T, h, W)))
        if self.mvl0.shape != (T, h, W, 3):
            raise ValueError("Sequence: mvl0 has shape %s, expected %s" % (self.mvl0.shape, (T, h, W, 3)))
        if self.gt is not None and self.gt.shape != (T, 4 * h, 4 * W):
            raise ValueError("Sequence: gt has shape %s, expected %s" % (self.gt.shape, (T, 4 * h, 4 * W)))

    @property
    def shape(self):
        return self.lr.shape

    @classmethod
    def from_directory(cls, lr_dir, side_dir, gt_dir=None, name=None):
        """Reads the reference's on-disk layout (test_LD_37.py:131-160): `lr_dir/<sorted frames>.png`,
        `side_dir/{part_m/%05d_M_mask.png, res/%05d_res.npy, unfiltered/%05d_unflt.png, mvl0/%05d_mvl0.npy}` and, optionally,
        `gt_dir/%05d.png` (metric/psnr_ssim.py:456-458).  Frame 0 has no side information; its slots repeat frame 1's."""
        import cv2
        files = sorted(f for f in os.listdir(lr_dir) if f.lower().endswith(".png"))
        if not files:
            raise FileNotFoundError("no PNG frames under %s" % lr_dir)

        def grey(path):
            img = cv2.imread(path, 0)
            if img is None:
                raise FileNotFoundError(path)
            return img

        lr = np.stack([grey(os.path.join(lr_dir, f)) for f in files])
        T = len(files)
        idx = ["%05d" % side_info_index(i) for i in range(T)]
        pm = np.stack([grey(os.path.join(side_dir, "part_m", k + "_M_mask.png")) for k in idx])
        unflt = np.stack([grey(os.path.join(side_dir, "unfiltered", k + "_unflt.png")) for k in idx])
        res = np.stack([np.load(os.path.join(side_dir, "res", k + "_res.npy"))[:, :, 0] for k in idx])
        mv = np.stack([np.load(os.path.join(side_dir, "mvl0", k + "_mvl0.npy")) for k in idx])
        gt = None
        if gt_dir is not None:
            gt = np.stack([grey(os.path.join(gt_dir, "%05d.png" % i)) for i in range(T)])
            gt = gt[:, :4 * lr.shape[1], :4 * lr.shape[2]]          # cal_psnr_ssim crops both to the common size (:460-466)
        return cls(lr, pm, res, unflt, mv, gt, name or os.path.basename(os.path.normpath(lr_dir)))


def padded_rows(h):
    """LR rows the model runs on: the reference appends 2 zero rows to 270-row frames (test_LD_37.py:24-26); in general
    the next multiple of 8 (CVSR_V8 needs H % 8 == 0)."""
    return (h + 7) // 8 * 8


def noise_for(seed, seq_ids, frame, H, W, device):
    """Six [S, 64, H, W] uniform tensors (neighbour order 0,1,2,4,5,6) keyed by (seed, sequence id, frame, neighbour)."""
    g = torch.Generator(device=device)
    out = []
    for nb in NEIGHBOURS:
        per_seq = []
        for sid in seq_ids:
            g.manual_seed(sharding.noise_key(seed, sid, frame, nb))
            per_seq.append(torch.rand((64, H, W), device=device, generator=g))
        out.append(torch.stack(per_seq).clamp_min_(1e-12))   # the reference redraws while any u == 0 (arch:2170-2171)
    return out


class _Staging:
    """One slot of pinned host staging + device landing buffers for the new frame of a step."""

    def __init__(self, S, h, H, W, mv_dtype, with_gt, device):
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()   # noqa: E731
        self.host = {"lr": pin((S, h, W), torch.uint8), "pm": pin((S, h, W), torch.uint8), "unflt": pin((S, h, W), torch.uint8),
                     "res": pin((S, h, W), torch.int16), "mv": pin((S, H, W, 3), mv_dtype)}
        self.host["mv"].zero_()                                    # padded rows: mv = refdist = 0 -> 0 / -0 = NaN -> flow 0
        if with_gt:
            self.host["gt"] = pin((S, 4 * h, 4 * W), torch.uint8)
        self.dev = {k: torch.empty_like(v, device=device) for k, v in self.host.items()}
        self.ready = torch.cuda.Event()      # recorded on the copy stream after the H2D copies
        self.consumed = torch.cuda.Event()   # recorded on the compute stream after the landing buffers were read
        self.used = False


class FrameDriver:
    """driver = FrameDriver(model); result = driver.run(sequences, seq_ids)

    `model` is a cdfo_b200.model.CVSR_V8 on a CUDA device in eval mode.  `run` processes a batch of sequences of equal
    (T, h, W) frame by frame and returns {"psnr": [..], "ssim": [..], "frames": T, "sums": float64 [S, 3] (device)}; `sink`,
    when given, is called as sink(batch_index, frame_index, uint8 ndarray [4h, 4W]) for every SR frame, one step late."""

    def __init__(self, model, seed=0, crop_border=4, graph=False):
        self.model, self.seed, self.crop_border, self.use_graph = model, int(seed), int(crop_border), bool(graph)
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise NotImplementedError("cdfo_b200.FrameDriver is CUDA-only (no CPU fallback)")
        self.copy_stream = torch.cuda.Stream(self.device)
        self._graphed = None

    # ------------------------------------------------------------------ host side of one step
    @staticmethod
    def _fill(slot, seqs, frame, mv_frame, h):
        for s, q in enumerate(seqs):
            k = side_info_index(frame)
            slot.host["lr"][s].copy_(torch.from_numpy(q.lr[frame]))
            slot.host["pm"][s].copy_(torch.from_numpy(q.pm[k]))
            slot.host["res"][s].copy_(torch.from_numpy(q.res[k]))
            slot.host["unflt"][s].copy_(torch.from_numpy(q.unflt[k]))
            slot.host["mv"][s, :h].copy_(torch.from_numpy(q.mvl0[side_info_index(mv_frame)]))
            if "gt" in slot.host:
                slot.host["gt"][s].copy_(torch.from_numpy(q.gt[mv_frame]))

    def _upload(self, slot):
        cs = self.copy_stream
        if slot.used:
            cs.wait_event(slot.consumed)
        with torch.cuda.stream(cs):
            for k, v in slot.host.items():
                slot.dev[k].copy_(v, non_blocking=True)
            slot.ready.record(cs)
        slot.used = True

    # ------------------------------------------------------------------ the loop
    @torch.no_grad()
    def run(self, sequences, seq_ids=None, sink=None):
        seqs = list(sequences)
        S = len(seqs)
        if S == 0:
            raise ValueError("FrameDriver.run: no sequences")
        T, h, W = seqs[0].shape
        if any(q.shape != (T, h, W) for q in seqs):
            raise ValueError("FrameDriver.run: the sequences of one batch must share (frames, rows, columns)")
        if W % 8:
            raise ValueError("FrameDriver.run: frame width must be a multiple of 8 (got %d)" % W)
        seq_ids = list(range(S)) if seq_ids is None else list(seq_ids)
        with_gt = all(q.gt is not None for q in seqs)
        # int32 staging if any sequence carries int32 MVs; int8 fields are widened by the staging copy itself (the caller's
        # Sequence objects are never modified)
        mv_dtype = torch.int32 if any(q.mvl0.dtype == np.int32 for q in seqs) else torch.int8
        dev, H = self.device, padded_rows(h)
        main = torch.cuda.current_stream(dev)
        slots = [_Staging(S, h, H, W, mv_dtype, with_gt, dev) for _ in range(2)]
        out_u8 = [torch.empty((S, 4 * h, 4 * W), dtype=torch.uint8, device=dev) for _ in range(2)]
        out_host = [torch.empty((S, 4 * h, 4 * W), dtype=torch.uint8).pin_memory() for _ in range(2)]
        out_done = [torch.cuda.Event() for _ in range(2)]
        out_ready = [torch.cuda.Event() for _ in range(2)]
        sums = torch.zeros((S, 3), dtype=torch.float64, device=dev)

        # ---- first window: frames o_list(0) of every plane (test_LD_37.py:143-155), converted on the device
        o0 = generate_input_index(0, 7, T - 1)

        def window_of(get):
            host = torch.from_numpy(np.stack([np.stack([get(q, j) for j in o0]) for q in seqs]))
            return metrics.planes_to_unit(host.to(dev), H).unsqueeze(2)          # [S, 7, 1, H, W]

        win = {"x": window_of(lambda q, j: q.lr[j]), "pms": window_of(lambda q, j: q.pm[side_info_index(j)]),
               "rms": window_of(lambda q, j: q.res[side_info_index(j)]), "ufs": window_of(lambda q, j: q.unflt[side_info_index(j)])}
        self._fill(slots[0], seqs, o0[-1], 0, h)       # step 0 needs only the MV field (+ GT) from its slot
        self._upload(slots[0])
        l1, pending = None, None
        ring_before = self.model.feature_ring
        self.model.feature_ring = not self.use_graph      # eager steps keep the window's L1 features in a ring (model.FeatureRing)
        try:
            for i in range(T):
                slot = slots[i % 2]
                if i + 1 < T:                               # stage step i+1 while step i computes
                    nxt = slots[(i + 1) % 2]
                    if nxt.used:
                        nxt.ready.synchronize()             # its previous H2D finished: the pinned buffers may be rewritten
                    self._fill(nxt, seqs, new_frame_of_step(i + 1, T), i + 1, h)
                    self._upload(nxt)
                main.wait_event(slot.ready)
                if i > 0:
                    new = {"x": metrics.planes_to_unit(slot.dev["lr"], H), "pms": metrics.planes_to_unit(slot.dev["pm"], H),
                           "rms": metrics.planes_to_unit(slot.dev["res"], H), "ufs": metrics.planes_to_unit(slot.dev["unflt"], H)}
                    for k in win:
                        win[k] = torch.cat([win[k][:, 1:], new[k].view(S, 1, 1, H, W)], 1)
                mvs = torch.cat([priors.mv2mvs(slot.dev["mv"][s]) for s in range(S)], 0)
                priors.modify_mv_for_end_frames(i, mvs, T)
                noise = noise_for(self.seed, seq_ids, i, H, W, dev)
                if l1 is None:
                    sr, l1 = self.model(win["x"], None, mvs, win["pms"], win["rms"], win["ufs"], None, noise=noise)
                elif self.use_graph:
                    if self._graphed is None or self._graphed.static_in[0].shape != win["x"].shape:
                        self._graphed = GraphedStep(self.model, win["x"], mvs, win["pms"], win["rms"], win["ufs"], l1, noise)
                    sr, l1 = self._graphed(win["x"], mvs, win["pms"], win["rms"], win["ufs"], l1, noise)
                else:
                    sr, l1 = self.model(win["x"], None, mvs, win["pms"], win["rms"], win["ufs"], l1, noise=noise)
                o = i % 2
                if i >= 2 and sink is not None:
                    main.wait_event(out_done[o])                      # the D2H of step i-2 has left this buffer
                metrics.sr_to_u8(sr.reshape(S, 4 * H, 4 * W), 4 * h, out=out_u8[o])
                if with_gt:
                    metrics.psnr_ssim(out_u8[o], slot.dev["gt"], self.crop_border, accum=sums)
                slot.consumed.record(main)
                out_ready[o].record(main)
                if sink is not None:
                    cs = self.copy_stream
                    cs.wait_event(out_ready[o])
                    with torch.cuda.stream(cs):
                        out_host[o].copy_(out_u8[o], non_blocking=True)
                        out_done[o].record(cs)
                    if pending is not None:                 # deliver the previous frame while this one computes
                        self._deliver(sink, pending, out_host, out_done)
                    pending = (i, o, S)
            if sink is not None and pending is not None:
                self._deliver(sink, pending, out_host, out_done)
            torch.cuda.current_stream(dev).synchronize()
        finally:
            self.model.feature_ring = ring_before       # also on an exception mid-sequence: later plain callers must not get ring handles
        res = {"frames": T, "sums": sums, "psnr": None, "ssim": None}
        if with_gt:
            host = sums.cpu()
            res["psnr"] = (host[:, 0] / host[:, 2]).tolist()       # mean of the per-frame values (metric/psnr_ssim.py:477-481)
            res["ssim"] = (host[:, 1] / host[:, 2]).tolist()
        return res

    @staticmethod
    def _deliver(sink, pending, out_host, out_done):
        i, o, S = pending
        out_done[o].synchronize()
        for s in range(S):
            sink(s, i, out_host[o][s].numpy().copy())


def run_sharded(model, sequences, batch=4, seed=0, crop_border=4, graph=False, sink=None, group=None):
    """Evaluates `sequences` (the full, rank-independent list) over all ranks of the process group: rank r owns a contiguous
    block of sequences (sharding.sequence_shard), processes it in batches of `batch`, and one all_reduce of the [n_seq, 3]
    fp64 sums (sharding.gather_metrics) gives every rank the per-sequence PSNR / SSIM.  No collective on the data path."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    n = len(sequences)
    own = sharding.sequence_shard(n, world, rank)
    drv = FrameDriver(model, seed=seed, crop_border=crop_border, graph=graph)
    sums = torch.zeros((n, 3), dtype=torch.float64, device=drv.device)
    for b0 in range(0, len(own), batch):
        ids = own[b0:b0 + batch]
        wrapped = None if sink is None else (lambda s, i, img, ids=ids: sink(ids[s], i, img))
        r = drv.run([sequences[k] for k in ids], seq_ids=ids, sink=wrapped)
        sums[ids[0]:ids[-1] + 1] = r["sums"]
    sums = sharding.gather_metrics(sums, group)
    host = sums.cpu()
    cnt = host[:, 2].clamp_min(1.0)
    return {"psnr": (host[:, 0] / cnt).tolist(), "ssim": (host[:, 1] / cnt).tolist(), "frames": host[:, 2].tolist(), "sums": sums}
