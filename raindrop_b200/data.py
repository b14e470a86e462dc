"""Input pipeline of code/Raindrop.py / code/utils_rd.py with the tensors resident on the device.

What the reference does on the host for every run (and partly for every step), and where it lives here:

  reference (file:line)                                    here
  -------------------------------------------------------  ---------------------------------------------------
  np.load('.../PTdict_list.npy', allow_pickle=True)        `load_ptdict_list` / `load_array_dataset` (host, once)
    (P12data/process_scripts/IrregularSampling.py:69-86,
     code/utils_rd.py:23-36)
  getStats (code/utils_rd.py:149-161)                      `feature_stats`      -> rd_feature_stats
  mask_normalize + concat(mask) + permute(1,0,2)           `tensorize_normalize[_other]` -> rd_mask_normalize
    (code/utils_rd.py:164-175,221-257, Raindrop.py:233)        (one kernel, written in the [T, n, 2F] training layout)
  mask_normalize_static (code/utils_rd.py:203-218)          same function (elementwise, torch on device)
  leave-sensors-out masks (code/Raindrop.py:214-231)        `remove_features_`   -> rd_zero_features
  balanced batches, strategy 2 / 3 (Raindrop.py:261-309)    `EpochSampler` (index matrix of a whole epoch, uploaded once)
  P[:, idx].cuda() ... lengths (Raindrop.py:311-317)        `DeviceDataset.fill` -> rd_assemble_batch (one launch)

Host random choices (np.random.shuffle / np.random.choice) are made with numpy in the SAME call order as the
reference, so with the same seed the batches and the removed sensors are the same ones.
"""
import numpy as np
import torch

from . import lib as L


# ---- on-disk formats ------------------------------------------------------------------------------
def load_ptdict_list(path):
    """P12 / P19 / eICU: a pickled list of per-patient dicts {'arr': [T, F], 'time': [T, 1] (minutes),
    'extended_static': [D], ...} (P12data/process_scripts/IrregularSampling.py:69-86).
    Returns raw float32 arrays P [n, T, F], minutes [n, T], static [n, D]."""
    lst = np.load(path, allow_pickle=True)
    n = len(lst)
    T, F = lst[0]["arr"].shape
    D = len(lst[0]["extended_static"])
    P = np.zeros((n, T, F), dtype=np.float32)
    minutes = np.zeros((n, T), dtype=np.float32)
    static = np.zeros((n, D), dtype=np.float32)
    for i, d in enumerate(lst):
        P[i] = d["arr"]
        minutes[i] = np.asarray(d["time"]).reshape(-1)
        static[i] = d["extended_static"]
    return P, minutes, static


def load_array_dataset(path):
    """PAM: `PTdict_list.npy` is a plain array [n, T, F]; timestamps are linspace(0, T, T) minutes
    (tensorize_normalize_other, code/utils_rd.py:243-257)."""
    P = np.asarray(np.load(path, allow_pickle=True), dtype=np.float32)
    n, T, _ = P.shape
    minutes = np.broadcast_to(torch.linspace(0, T, T).numpy()[None, :], (n, T)).astype(np.float32).copy()
    return P, minutes, None


def load_split(split_path):
    """(idx_train, idx_val, idx_test) as saved by the reference's split scripts (code/utils_rd.py:10-20)."""
    a = np.load(split_path, allow_pickle=True)
    return a[0], a[1], a[2]


# ---- statistics / normalisation on the device -----------------------------------------------------------
def _dev(x, device, dtype=torch.float32):
    return torch.as_tensor(x).to(device=device, dtype=dtype).contiguous()


def feature_stats(P_raw):
    """getStats (code/utils_rd.py:149-161): per-feature mean / std over the observed (> 0) entries of P_raw [n, T, F]
    (device tensor).  Returns (mean [F], std [F]) on the device."""
    lib = L.load()
    n, T, F = P_raw.shape
    mean = torch.empty(F, dtype=torch.float32, device=P_raw.device)
    std = torch.empty_like(mean)
    sc = torch.empty(lib.rd_feature_stats_scratch_bytes(n, T, F), dtype=torch.uint8, device=P_raw.device)
    L.check(lib.rd_feature_stats(P_raw.data_ptr(), n, T, F, mean.data_ptr(), std.data_ptr(), sc.data_ptr(),
                                 L.stream_ptr(P_raw.device)), "rd_feature_stats")
    return mean, std


def get_stats_static(P_static, dataset="P12"):
    """getStats_static (code/utils_rd.py:178-200).  The reference tests `bool_categorical == 0` on the LIST, which is
    never true, so it always returns zeros / ones; reproduced as is."""
    S = P_static.shape[1]
    return torch.zeros(S, device=P_static.device), torch.ones(S, device=P_static.device)


def mask_normalize_static(P_static, ms, ss):
    """mask_normalize_static (code/utils_rd.py:203-218): normalise, then zero everything <= 0."""
    z = (P_static - ms[None, :]) / (ss[None, :] + 1e-18)
    return torch.where(z <= 0, torch.zeros_like(z), z)


def tensorize_normalize(P_raw, minutes, static, y, mf, stdf, ms=None, ss=None, device="cuda"):
    """tensorize_normalize / tensorize_normalize_other (code/utils_rd.py:221-257) + the permutes of
    code/Raindrop.py:233-239, on the device.  Inputs: raw arrays [n, T, F], [n, T], [n, D] | None, labels [n, 1] | [n].
    Returns (P [T, n, 2F], Pstatic [n, D] | None, Ptime [T, n], y [n] int64), all device tensors."""
    lib = L.load()
    P_raw = _dev(P_raw, device)
    minutes = _dev(minutes, device)
    n, T, F = P_raw.shape
    out = torch.empty(T, n, 2 * F, dtype=torch.float32, device=device)
    times = torch.empty(T, n, dtype=torch.float32, device=device)
    L.check(lib.rd_mask_normalize(P_raw.data_ptr(), mf.data_ptr(), stdf.data_ptr(), n, T, F, out.data_ptr(),
                                  minutes.data_ptr(), times.data_ptr(), L.stream_ptr(out.device)), "rd_mask_normalize")
    st = None
    if static is not None:
        st = _dev(static, device)
        if ms is None:
            ms, ss = get_stats_static(st)
        st = mask_normalize_static(st, ms, ss)
    yt = torch.as_tensor(np.asarray(y)).reshape(len(y), -1)[:, 0].to(device=device, dtype=torch.int64)
    return out, st, times, yt


# ---- leave-sensors-out (code/Raindrop.py:214-231) -------------------------------------------------------
def removal_indices(n_samples, num_features, missing_ratio, level="sample", density_scores=None):
    """Sensor indices the reference would zero: per sample `np.random.choice(F, k, replace=False)` in sample order
    (level 'sample'), or the first k of the information-gain ranking (level 'set').  Uses the global numpy RNG like the
    reference, so `np.random.seed(...)` reproduces its choice."""
    k = round(missing_ratio * num_features)
    if level == "sample":
        return np.stack([np.random.choice(num_features, k, replace=False) for _ in range(n_samples)]).astype(np.int64)
    if density_scores is None:
        raise ValueError("feature_removal_level 'set' needs the density-score ranking (IG_density_scores_*.npy[:, 0])")
    return np.asarray(density_scores[:k]).astype(np.int64)


def remove_features_(P, idx):
    """In place on a device tensor P [T, B, 2F]: zero the value columns idx ([B, k] per sample or [k] for the set)."""
    lib = L.load()
    idx_t = torch.as_tensor(idx).to(device=P.device, dtype=torch.int64).contiguous()
    per_sample = 1 if idx_t.dim() == 2 else 0
    K = idx_t.shape[-1]
    T, B, W = P.shape
    if per_sample and idx_t.shape[0] != B:
        raise ValueError("per-sample removal indices must be [B, k]")
    L.check(lib.rd_zero_features(P.data_ptr(), T, B, W, idx_t.data_ptr(), K, per_sample, L.stream_ptr(P.device)),
            "rd_zero_features")
    return P


# ---- batch index generation (code/Raindrop.py:261-309) --------------------------------------------------
class EpochSampler:
    """Index matrix [n_batches, batch_size] of one epoch, generated on the host exactly like the reference and
    uploaded ONCE per epoch (instead of one host->device batch copy per step).

    strategy 2 (P12 / P19 / eICU): minority class upsampled 3x, both index lists shuffled per epoch
        (`np.random.shuffle(expanded_idx_1)` then `np.random.shuffle(idx_0)`), each batch = B/2 negatives + B/2 positives;
        n_batches = min(n0 // (B/2), 3 n1 // (B/2)).
    strategy 3 (PAM): 30 batches of `np.random.choice(n, B, replace=False)`."""

    def __init__(self, y, batch_size=128, strategy=2, device="cuda"):
        y = np.asarray(y).reshape(len(y), -1)[:, 0]
        self.B, self.strategy, self.device, self.n = int(batch_size), strategy, device, len(y)
        self.idx_0 = np.where(y == 0)[0]
        idx_1 = np.where(y == 1)[0]
        self.expanded_idx_1 = np.concatenate([idx_1, idx_1, idx_1], axis=0)
        if strategy == 2:
            half = self.B // 2
            self.n_batches = int(min(len(self.idx_0) // half, len(self.expanded_idx_1) // half))
        elif strategy == 3:
            self.n_batches = 30
        else:
            raise ValueError("strategy must be 2 or 3 (the ones code/Raindrop.py:264-267 selects)")

    def epoch(self):
        """-> int64 device tensor [n_batches, B]"""
        if self.strategy == 2:
            half = self.B // 2
            np.random.shuffle(self.expanded_idx_1)
            np.random.shuffle(self.idx_0)
            rows = [np.concatenate([self.idx_0[n * half:(n + 1) * half], self.expanded_idx_1[n * half:(n + 1) * half]])
                    for n in range(self.n_batches)]
        else:
            rows = [np.random.choice(list(range(self.n)), size=self.B, replace=False) for _ in range(self.n_batches)]
        return torch.as_tensor(np.stack(rows).astype(np.int64)).to(self.device)


class DeviceDataset:
    """Training / evaluation tensors of code/Raindrop.py:233-239 resident in HBM.  `fill` assembles one batch --
    gather of the B samples, their times, statics, labels AND lengths = #(t > 0) (Raindrop.py:317) -- in ONE kernel
    launch straight into a TrainStep's static buffers (or any buffers with those attributes), so a step moves no batch
    data over PCIe.  `removed`: optional leave-sensors-out indices applied to the assembled batch."""

    def __init__(self, P, Pstatic, Ptime, y, device="cuda"):
        self.P = _dev(P, device)                                  # [T, n, 2F]
        self.Ptime = _dev(Ptime, device)                          # [T, n]
        self.Pstatic = None if Pstatic is None else _dev(Pstatic, device)
        self.y = None if y is None else torch.as_tensor(y).to(device=device, dtype=torch.int64).reshape(-1).contiguous()
        self.T, self.n, self.width = self.P.shape
        self.lib = L.load()

    def fill(self, step, idx, removed=None):
        idx = idx.to(device=self.P.device, dtype=torch.int64, non_blocking=True).contiguous()
        B = idx.numel()
        st = getattr(step, "static", None)
        if (st is None) != (self.Pstatic is None):
            st = None
        ds = 0 if st is None else self.Pstatic.shape[1]
        L.check(self.lib.rd_assemble_batch(self.P.data_ptr(), self.Ptime.data_ptr(), L.ptr(self.Pstatic if st is not None else None),
                                           L.ptr(self.y), idx.data_ptr(), self.T, self.n, self.width, ds, B,
                                           step.src.data_ptr(), step.times.data_ptr(), L.ptr(st), L.ptr(step.y if self.y is not None else None),
                                           step.lengths.data_ptr(), L.stream_ptr(self.P.device)), "rd_assemble_batch")
        if removed is not None:
            remove_features_(step.src, removed)


class BatchBuffers:
    """Plain holder with the attributes `DeviceDataset.fill` writes (for evaluation without a TrainStep)."""

    def __init__(self, T, B, width, d_static, device="cuda"):
        self.src = torch.empty(T, B, width, dtype=torch.float32, device=device)
        self.times = torch.empty(T, B, dtype=torch.float32, device=device)
        self.static = torch.empty(B, d_static, dtype=torch.float32, device=device) if d_static else None
        self.y = torch.empty(B, dtype=torch.int64, device=device)
        self.lengths = torch.empty(B, dtype=torch.int64, device=device)
