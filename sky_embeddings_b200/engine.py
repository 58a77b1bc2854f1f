"""Host-side handle over the C ABI: a device-resident embedding bank and exact top-k search.

PyTorch is used for device memory, streams and (in distributed.py) NCCL plumbing only; all
arithmetic of the search path runs in libskysearch.so.  Reference path being replaced:
/root/reference/utils/similarity.py (mae_simsearch :37-132, compute_similarity :214-268,
update_best_scores :18-35).
"""
from __future__ import annotations

import ctypes as C
import ctypes as C_

import torch

from . import _lib as L

_TOKEN_MODES = {"all": L.TOK_ALL, "cls": L.TOK_CLS, "patches": L.TOK_PATCHES, "maxpool": L.TOK_MAXPOOL}
_DTYPES = {"fp32": L.F32, "f32": L.F32, "float32": L.F32, torch.float32: L.F32,
           "bf16": L.BF16, "bfloat16": L.BF16, torch.bfloat16: L.BF16}


def token_mode_of(max_pool=False, cls_token=False):
    """The reference's two switches (utils/similarity.py:55-63) -> one token mode."""
    if cls_token:
        return "cls"
    return "maxpool" if max_pool else "patches"


def tokens_kept(tokens, mode, num_extra_tokens=1):
    return {"all": tokens, "cls": 1, "patches": tokens - num_extra_tokens, "maxpool": 1}[mode]


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _src(t):
    if t.dtype == torch.float32:
        return L.F32
    if t.dtype == torch.bfloat16:
        return L.BF16
    raise TypeError(f"unsupported source dtype {t.dtype} (float32 or bfloat16)")


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class Bank:
    """N items x L tokens x D features, normalised once with the first-batch statistics
    (utils/similarity.py:98-102) and stored as bf16 or fp32 rows in HBM."""

    def __init__(self, n_items, L_tokens, D, dtype="bf16", device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("sky_embeddings_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = L.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_items, self.L, self.D = int(n_items), int(L_tokens), int(D)
        self.dtype = _DTYPES[dtype]
        h = C.c_void_p()
        L.check(self.lib.sky_bank_create(C.byref(h), self.device.index, self.n_items, self.L, self.D, self.dtype))
        self._h = h
        self.finalized = False
        self.has_norm = False

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.sky_bank_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- building ---------------------------------------------------------------------------
    def _prep(self, x):
        if x.device != self.device:
            x = x.to(self.device, non_blocking=True)
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        if x.dim() == 2:
            x = x.unsqueeze(1)
        if x.dim() != 3 or x.shape[-1] != self.D:
            raise ValueError(f"expected [items, tokens, {self.D}] latents, got {tuple(x.shape)}")
        return x

    def fit_norm(self, first_batch, token_mode="all", num_extra_tokens=1):
        x = self._prep(first_batch)
        L.check(self.lib.sky_bank_fit_norm(self._h, _ptr(x), _src(x), x.shape[0], x.shape[1],
                                           _TOKEN_MODES[token_mode], num_extra_tokens, _stream(self.device)))
        self.has_norm = True
        return self

    def set_norm(self, mu, sigma):
        mu = mu.to(self.device, torch.float32).contiguous()
        sigma = sigma.to(self.device, torch.float32).contiguous()
        L.check(self.lib.sky_bank_set_norm(self._h, _ptr(mu), _ptr(sigma), _stream(self.device)))
        self.has_norm = True
        return self

    def norm(self):
        mu = torch.empty(self.D, device=self.device, dtype=torch.float32)
        sigma = torch.empty_like(mu)
        L.check(self.lib.sky_bank_get_norm(self._h, _ptr(mu), _ptr(sigma), _stream(self.device)))
        return mu, sigma

    def upload(self, latents, item0=0, token_mode="all", num_extra_tokens=1):
        x = self._prep(latents)
        L.check(self.lib.sky_bank_upload(self._h, _ptr(x), _src(x), int(item0), x.shape[0], x.shape[1],
                                         _TOKEN_MODES[token_mode], num_extra_tokens, _stream(self.device)))
        self.finalized = False
        return self

    def finalize(self):
        L.check(self.lib.sky_bank_finalize(self._h, _stream(self.device)))
        self.finalized = True
        return self

    def resize(self, n_items):
        """Change the active item count within the capacity (streaming callers reuse one bank)."""
        L.check(self.lib.sky_bank_resize(self._h, int(n_items)))
        self.n_items = int(n_items)
        return self

    def profile(self, enable=True):
        """Time the scoring kernel of every search with CUDA events on its launch stream."""
        L.check(self.lib.sky_profile_enable(self._h, 1 if enable else 0))
        return self

    def profile_read(self, reset=True):
        """(number of scoring-kernel launches, their total duration in ms) since the last reset."""
        n, ms = C.c_int64(), C.c_double()
        L.check(self.lib.sky_profile_read(self._h, C.byref(n), C.byref(ms), 1 if reset else 0))
        return int(n.value), float(ms.value)

    def download(self, item0=0, n_items=None):
        n = self.n_items - item0 if n_items is None else n_items
        out = torch.empty((n, self.L, self.D), device=self.device, dtype=torch.float32)
        L.check(self.lib.sky_bank_download(self._h, int(item0), int(n), _ptr(out), _stream(self.device)))
        return out

    @classmethod
    def from_latents(cls, latents, norm_rows=None, token_mode="all", num_extra_tokens=1, dtype="bf16",
                     device=None, chunk_items=1 << 16):
        """Bank from encoder output [N, tokens, D]; norm_rows = the reference's batch_size (the first
        batch defines the normalisation, utils/similarity.py:98-100); None = no normalisation."""
        if latents.dim() == 2:
            latents = latents.unsqueeze(1)
        n, tokens, D = latents.shape
        bank = cls(n, tokens_kept(tokens, token_mode, num_extra_tokens), D, dtype, device)
        if norm_rows is not None:
            bank.fit_norm(latents[:norm_rows], token_mode, num_extra_tokens)
        for s in range(0, n, chunk_items):
            bank.upload(latents[s:s + chunk_items], s, token_mode, num_extra_tokens)
        return bank.finalize()

    # -- persisted bank cache --------------------------------------------------------------------
    def save(self, path, chunk_items=1 << 18):
        """Persist the resident bank (stored, normalised rows + the first-batch statistics) so that later runs
        skip the encoder and the ingest statistics: the reference re-encodes the whole bank for every target
        (utils/similarity.py:71-102).  bf16 banks are written as their 16-bit patterns (exact)."""
        import numpy as np
        rows = np.empty((self.n_items, self.L, self.D), dtype=np.uint16 if self.dtype == L.BF16 else np.float32)
        for s0 in range(0, self.n_items, chunk_items):
            x = self.download(s0, min(chunk_items, self.n_items - s0))
            if self.dtype == L.BF16:
                rows[s0:s0 + x.shape[0]] = x.to(torch.bfloat16).view(torch.int16).cpu().numpy().view(np.uint16)
            else:
                rows[s0:s0 + x.shape[0]] = x.cpu().numpy()
        mu, sigma = (t.cpu().numpy() for t in self.norm()) if self.has_norm else (np.zeros(0, np.float32),) * 2
        np.savez(path, rows=rows, mu=mu, sigma=sigma, meta=np.array([self.n_items, self.L, self.D, self.dtype], dtype=np.int64))

    @classmethod
    def load(cls, path, device=None, chunk_items=1 << 18):
        """Rebuild a bank saved with save(): rows are uploaded as stored (no re-normalisation), then the saved
        statistics are attached so that query_from_targets normalises target groups exactly as before."""
        import numpy as np
        z = np.load(path)
        n, Lt, D, dt = (int(v) for v in z["meta"])
        bank = cls(n, Lt, D, "bf16" if dt == L.BF16 else "fp32", device)
        rows = z["rows"]
        for s0 in range(0, n, chunk_items):
            part = torch.from_numpy(np.ascontiguousarray(rows[s0:s0 + chunk_items]))
            if dt == L.BF16:
                part = part.view(torch.int16).view(torch.bfloat16)
            bank.upload(part, s0, "all", 0)
        if z["mu"].size:
            bank.set_norm(torch.from_numpy(z["mu"]), torch.from_numpy(z["sigma"]))
        return bank.finalize()

    # -- queries ----------------------------------------------------------------------------
    def query_from_targets(self, targets, use_weights=True):
        """determine_target_features (utils/similarity.py:134-147) on the normalised target group.
        targets: [T, L_t, D] or [T_rows, D], already token-selected."""
        x = targets.to(self.device, torch.float32).reshape(-1, self.D).contiguous()
        t = torch.empty(self.D, device=self.device, dtype=torch.float32)
        w = torch.empty_like(t)
        L.check(self.lib.sky_query_from_targets(self._h, _ptr(x), x.shape[0], self.D, 1 if use_weights else 0,
                                                _ptr(t), _ptr(w), _stream(self.device)))
        return t, w

    def _qprep(self, t, w):
        t = t.to(self.device, torch.float32)
        if t.dim() == 1:
            t = t.unsqueeze(0)
        t = t.contiguous()
        if t.shape[1] != self.D:
            raise ValueError(f"queries must be [Q, {self.D}]")
        if w is not None:
            w = w.to(self.device, torch.float32)
            if w.dim() == 1:
                w = w.unsqueeze(0)
            w = w.expand_as(t).contiguous()
        return t, w

    def search(self, t, w=None, k=100, metric="cosine", combine="min", n_top_sims=None, path="auto",
               idx_offset=0, out_scores=None, out_idx=None):
        """Exact top-k of every query over the bank.  Returns (scores [Q,k] f32, idx [Q,k] i64).
        out_scores / out_idx: optional preallocated contiguous device outputs."""
        if metric not in L.METRICS:
            # the reference dies with UnboundLocalError here (utils/similarity.py:250-259)
            raise ValueError(f"unknown metric {metric!r}: expected 'cosine', 'MSE' or 'MAE'")
        if combine not in L.COMBINES:
            raise ValueError(f"unknown combine {combine!r}: expected 'mean', 'min' or 'max'")
        t, w = self._qprep(t, w)
        Q = t.shape[0]
        scores = torch.empty((Q, k), device=self.device, dtype=torch.float32) if out_scores is None else out_scores
        idx = torch.empty((Q, k), device=self.device, dtype=torch.int64) if out_idx is None else out_idx
        if scores.shape != (Q, k) or idx.shape != (Q, k) or not (scores.is_contiguous() and idx.is_contiguous()):
            raise ValueError("out_scores / out_idx must be contiguous [Q, k] tensors")
        L.check(self.lib.sky_search(self._h, _ptr(t), _ptr(w) if w is not None else None, Q, L.METRICS[metric],
                                    L.COMBINES[combine], int(n_top_sims or 0), int(k), int(idx_offset),
                                    _ptr(scores), _ptr(idx), L.PATHS[path], _stream(self.device)))
        return scores, idx

    def search_sharded(self, exchange_handle, t, w=None, k=100, metric="cosine", combine="min", n_top_sims=None,
                       path="auto", idx_offset=0, out_scores=None, out_idx=None):
        """search() over this rank's shard + the peer-memory candidate exchange in one call (sky_search_sharded): returns
        the GLOBAL top-k, identical on every rank.  exchange_handle: a connected sky_exchange (distributed.PeerExchange)."""
        if metric not in L.METRICS:
            raise ValueError(f"unknown metric {metric!r}: expected 'cosine', 'MSE' or 'MAE'")
        if combine not in L.COMBINES:
            raise ValueError(f"unknown combine {combine!r}: expected 'mean', 'min' or 'max'")
        t, w = self._qprep(t, w)
        Q = t.shape[0]
        scores = torch.empty((Q, k), device=self.device, dtype=torch.float32) if out_scores is None else out_scores
        idx = torch.empty((Q, k), device=self.device, dtype=torch.int64) if out_idx is None else out_idx
        L.check(self.lib.sky_search_sharded(self._h, exchange_handle, _ptr(t), _ptr(w) if w is not None else None, Q,
                                            L.METRICS[metric], L.COMBINES[combine], int(n_top_sims or 0), int(k),
                                            int(idx_offset), _ptr(scores), _ptr(idx), L.PATHS[path], _stream(self.device)))
        return scores, idx

    def search_host(self, t_host, w_host=None, k=100, metric="cosine", combine="min", n_top_sims=None,
                    path="auto", idx_offset=0, out_scores=None, out_idx=None):
        """Same as search() with HOST tensors in and out; the H2D / D2H copies happen inside the call."""
        t_host = t_host.contiguous()
        if t_host.dim() == 1:
            t_host = t_host.unsqueeze(0)
        Q = t_host.shape[0]
        if out_scores is None:
            out_scores = torch.empty((Q, k), dtype=torch.float32, pin_memory=True)
        if out_idx is None:
            out_idx = torch.empty((Q, k), dtype=torch.int64, pin_memory=True)
        L.check(self.lib.sky_search_host(self._h, _ptr(t_host), _ptr(w_host) if w_host is not None else None, Q,
                                         L.METRICS[metric], L.COMBINES[combine], int(n_top_sims or 0), int(k),
                                         int(idx_offset), _ptr(out_scores), _ptr(out_idx), L.PATHS[path],
                                         _stream(self.device)))
        return out_scores, out_idx

    def score(self, t, w=None, metric="cosine", combine="min", n_top_sims=None, item0=0, n_items=None):
        """compute_similarity (utils/similarity.py:214-268) of items [item0, item0+n) -> [Q, n] f32."""
        if metric not in L.METRICS:
            raise ValueError(f"unknown metric {metric!r}: expected 'cosine', 'MSE' or 'MAE'")
        if combine not in L.COMBINES:
            raise ValueError(f"unknown combine {combine!r}: expected 'mean', 'min' or 'max'")
        t, w = self._qprep(t, w)
        n = self.n_items - item0 if n_items is None else n_items
        out = torch.empty((t.shape[0], n), device=self.device, dtype=torch.float32)
        L.check(self.lib.sky_score(self._h, _ptr(t), _ptr(w) if w is not None else None, t.shape[0],
                                   L.METRICS[metric], L.COMBINES[combine], int(n_top_sims or 0), int(item0), int(n),
                                   _ptr(out), _stream(self.device)))
        return out


class PixelBank:
    """Raw cutouts [N, C, H, W] fp32 (NaN = missing pixel / band) resident in HBM, searched in pixel space
    with the NaN-aware masked MSE of BASELINE config 5:
    score = sum m (q - x)^2 / (sum m + 1e-5), m = ~isnan(q) & ~isnan(x) & qmask; lower is better."""

    def __init__(self, n_items, C, H, W, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("sky_embeddings_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = L.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_items, self.shape, self.D = int(n_items), (int(C), int(H), int(W)), int(C) * int(H) * int(W)
        h = C_.c_void_p()
        L.check(self.lib.sky_pixel_bank_create(C_.byref(h), self.device.index, self.n_items, *self.shape))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.sky_bank_destroy(self._h)
            self._h = C_.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, cutouts, item0=0):
        """cutouts [n, C, H, W] (or [n, D]) float32, on this device or on the host."""
        x = cutouts.to(torch.float32).contiguous()
        if x.is_cuda and x.device != self.device:
            x = x.to(self.device)
        if x[0].numel() != self.D:
            raise ValueError(f"expected cutouts of {self.shape}, got {tuple(x.shape[1:])}")
        L.check(self.lib.sky_pixel_bank_upload(self._h, _ptr(x), int(item0), x.shape[0], _stream(self.device)))
        if not x.is_cuda:
            torch.cuda.current_stream(self.device).synchronize()      # the host buffer may go away
        return self

    @classmethod
    def from_cutouts(cls, cutouts, device=None, chunk_items=4096):
        n = cutouts.shape[0]
        bank = cls(n, *cutouts.shape[1:], device=device)
        for s in range(0, n, chunk_items):
            bank.upload(cutouts[s:s + chunk_items], s)
        return bank

    def profile(self, enable=True):
        L.check(self.lib.sky_profile_enable(self._h, 1 if enable else 0))
        return self

    def profile_read(self, reset=True):
        n, ms = C_.c_int64(), C_.c_double()
        L.check(self.lib.sky_profile_read(self._h, C_.byref(n), C_.byref(ms), 1 if reset else 0))
        return int(n.value), float(ms.value)

    def _qprep(self, q, qmask):
        q = q.to(self.device, torch.float32).reshape(-1, self.D).contiguous() if q.dim() != 1 else \
            q.to(self.device, torch.float32).reshape(1, self.D).contiguous()
        if qmask is not None:
            qmask = (qmask.to(self.device) != 0).to(torch.uint8).reshape(-1, self.D)
            qmask = qmask.expand(q.shape[0], self.D).contiguous()
        return q, qmask

    def search(self, q, qmask=None, k=100, idx_offset=0):
        """q [Q, C, H, W] (NaN = missing), qmask same shape (non-zero = compare) or None.
        Returns (scores [Q, k] ascending, idx [Q, k] i64)."""
        q, qmask = self._qprep(q, qmask)
        Q = q.shape[0]
        scores = torch.empty((Q, k), device=self.device, dtype=torch.float32)
        idx = torch.empty((Q, k), device=self.device, dtype=torch.int64)
        L.check(self.lib.sky_search_pixels(self._h, _ptr(q), _ptr(qmask) if qmask is not None else None, Q, int(k),
                                           int(idx_offset), _ptr(scores), _ptr(idx), _stream(self.device)))
        return scores, idx

    def score(self, q, qmask=None, item0=0, n_items=None):
        q, qmask = self._qprep(q, qmask)
        n = self.n_items - item0 if n_items is None else n_items
        out = torch.empty((q.shape[0], n), device=self.device, dtype=torch.float32)
        L.check(self.lib.sky_score_pixels(self._h, _ptr(q), _ptr(qmask) if qmask is not None else None, q.shape[0],
                                          int(item0), int(n), _ptr(out), _stream(self.device)))
        return out


def merge_candidates(scores, idx, k_out, metric):
    """Merge R candidate lists per query (scores/idx [R, Q, k_in]) into the best-first top-k_out.
    The device merge after the all-gather, and update_best_scores' running merge
    (utils/similarity.py:18-35)."""
    lib = L.load()
    if not scores.is_cuda:
        raise RuntimeError("merge_candidates needs CUDA tensors; there is no CPU fallback")
    scores = scores.to(torch.float32).contiguous()
    idx = idx.to(torch.int64).contiguous()
    R, Q, k_in = scores.shape
    out_s = torch.empty((Q, k_out), device=scores.device, dtype=torch.float32)
    out_i = torch.empty((Q, k_out), device=scores.device, dtype=torch.int64)
    L.check(lib.sky_merge_candidates(_ptr(scores), _ptr(idx), R, Q, k_in, int(k_out), L.METRICS[metric],
                                     _ptr(out_s), _ptr(out_i), scores.device.index, _stream(scores.device)))
    return out_s, out_i
