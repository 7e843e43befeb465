"""Host-side engine of the multi-view denoising hot path.

A denoising step (GridAttn over N views -> UNet over N (x2 under CFG) view latents -> CFG combine -> DDIM
update) is compiled ONCE per shape into a `Program`: a flat list of bound C-ABI calls (ops.BoundCall) over
pre-allocated HBM buffers.  The program is replayed every step, either call by call or from a captured CUDA
graph; nothing is allocated, synchronised or copied to the host inside it.

Data layout in HBM: activations are "rows x channels" (NHWC flattened), the residual stream is
fp32, every tensor-core operand is fp16; weights are packed once per state-dict version into K-major fp16
matrices (conv kernels as [C_out, (ky,kx,c)], q/k/v fused, GEGLU value/gate rows interleaved per 128-column tile).

Reference call sites each emitter replaces are cited in its docstring (paths under the reference root).
"""
import json
import math
import os

import torch

from .ops import ACT_GEGLU, ACT_GELU, ACT_NONE, ACT_SILU, gemm_signature

_TUNING = None


def gemm_tuning():
    """signature -> (tile_n, split_k, cta_pair), measured on a B200 by tools/tune_gemm.py (empty when the file is absent or
    MVD_GEMM_NO_TUNING is set: the library's own heuristics then decide)."""
    global _TUNING
    if _TUNING is None:
        _TUNING = {}
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_tuning.json")
        if os.path.exists(path) and not os.environ.get("MVD_GEMM_NO_TUNING"):
            with open(path) as f:
                _TUNING = {k: tuple(v) for k, v in json.load(f).get("choices", {}).items()}
    return _TUNING

NUM_SMS = 148
Z_CH = 256        # GridAttn z_embedder width (mvdfusion/view_attn_efficient2.py:151)
TOKEN_LD = 736    # 723-d GridAttn token rounded up to a multiple of 16
CTX_DIM = 768
SPLITK_WS_BYTES = 64 << 20
GEGLU_TILE = 256  # value | gate column interleave period of the packed GEGLU weights (= the GEMM tile width)


def _round_up(x, m):
    return (x + m - 1) // m * m


# ------------------------------------------------------------------------------------------------ weights
class PackedWeights:
    """Kernel-side weight cache derived from a reference-named state dict (SURVEY.md §5 checkpoint contract:
    parameter names / shapes stay the reference's; packed copies are rebuilt after load_state_dict)."""

    def __init__(self, state_dict, ops, prefix=""):
        self.sd = state_dict
        self.ops = ops
        self.prefix = prefix
        self.cache = {}

    def _full(self, key):
        k = self.prefix + key
        return k[1:] if k.startswith(".") else k  # module-level plans use an empty layer prefix

    def has(self, key):
        return self._full(key) in self.sd

    def raw(self, key):
        return self.sd[self._full(key)].detach()

    def _put(self, tag, key, fn):
        k = (tag, key)
        if k not in self.cache:
            self.cache[k] = fn().contiguous().to(self.ops.device)
        return self.cache[k]

    def f32(self, key):
        return self._put("f32", key, lambda: self.raw(key).float().reshape(-1))

    def f32_sum(self, key_a, key_b):
        return self._put("f32sum", (key_a, key_b), lambda: (self.raw(key_a).float() + self.raw(key_b).float()).reshape(-1))

    @staticmethod
    def _pad_k(w):
        n, k = w.shape
        k8 = _round_up(k, 8)
        if k8 != k:
            w = torch.cat([w, w.new_zeros(n, k8 - k)], dim=1)
        return w

    def lin(self, key, k_pad=None):
        """nn.Linear / 1x1-conv weight [N, K(,1,1)] -> fp16 [N, K8]"""
        def f():
            w = self.raw(key).float()
            w = w.reshape(w.shape[0], -1)
            if k_pad is not None and k_pad > w.shape[1]:
                w = torch.cat([w, w.new_zeros(w.shape[0], k_pad - w.shape[1])], dim=1)
            return self._pad_k(w).half()
        return self._put("lin", (key, k_pad), f)

    @staticmethod
    def _head_order(heads, hd):
        """row permutation of a timm qkv Linear ([q | k | v] x heads x hd) into head-major [q_h | k_h | v_h] blocks
        (mvd_gridattn_dit_f16 computes one head's q, k, v as one 96-column product)"""
        C = heads * hd
        idx = torch.arange(3 * C).reshape(3, heads, hd).permute(1, 0, 2).reshape(-1)
        return idx

    def qkv_heads(self, key, heads, hd):
        return self._put("qkv_heads", key, lambda: self.raw(key).float()[self._head_order(heads, hd)].half())

    def f32_heads(self, key, heads, hd):
        return self._put("f32_heads", key, lambda: self.raw(key).float().reshape(-1)[self._head_order(heads, hd)])

    def conv3(self, key, c_pad=None):
        """conv3x3 weight [Cout, Cin, 3, 3] -> fp16 [Cout, 9*Cin_pad], k = (ky*3 + kx)*Cin_pad + c"""
        def f():
            w = self.raw(key).float()
            co, ci = w.shape[0], w.shape[1]
            cp = c_pad if c_pad is not None else ci
            w = w.permute(0, 2, 3, 1)  # [Cout, ky, kx, Cin]
            if cp != ci:
                w = torch.cat([w, w.new_zeros(co, 3, 3, cp - ci)], dim=3)
            return w.reshape(co, 9 * cp).half()
        return self._put("conv3", (key, c_pad), f)

    def conv3_up2(self, key):
        """conv3x3 weight [Cout, Cin, 3, 3] behind a nearest x2 upsample -> the four 2 x 2 phase convolutions of the source image
        (include/mvd_b200.h, conv_up2): fp16 [4*Cout, 4*Cin], row (2 py + px) * Cout + n, column (2 a + b) * Cin + c, each entry the fp32 sum
        of the 3 x 3 taps that land on source tap (a, b) for output phase (py, px)."""
        def f():
            w = self.raw(key).double().permute(0, 2, 3, 1)  # [Cout, ky, kx, Cin]
            sets = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}   # phase -> taps summed onto source tap 0 / 1
            blocks = []
            for py in range(2):
                for px in range(2):
                    taps = []
                    for a in range(2):
                        for b in range(2):
                            taps.append(sum(w[:, ky, kx, :] for ky in sets[py][a] for kx in sets[px][b]))
                    blocks.append(torch.cat(taps, dim=1))
            return torch.cat(blocks, dim=0).float().half()
        return self._put("conv3_up2", key, f)

    def f32_tiled(self, key, reps):
        return self._put("f32t", (key, reps), lambda: self.raw(key).float().reshape(-1).repeat(reps))

    @staticmethod
    def _hilo(w):
        """fp32 -> (fp16 hi, fp16 lo) with hi + lo == w to ~2^-22 relative (include/mvd_b200.h, mvd_gemm_args.hilo)"""
        hi = w.half()
        return hi, (w - hi.float()).half()

    def lin_hilo(self, key):
        """nn.Linear / 1x1-conv weight [N, K] -> fp16 [N, 2K] = [W_hi | W_lo] (K a multiple of 64)"""
        def f():
            w = self.raw(key).float()
            w = w.reshape(w.shape[0], -1)
            assert w.shape[1] % 64 == 0
            return torch.cat(self._hilo(w), dim=1)
        return self._put("lin_hilo", key, f)

    def conv3_hilo(self, key):
        """conv3x3 weight [Cout, Cin, 3, 3] -> fp16 [Cout, 2*9*Cin] = [W_hi | W_lo], each half in the implicit-GEMM k order"""
        def f():
            w = self.raw(key).float()
            co, ci = w.shape[0], w.shape[1]
            assert ci % 64 == 0
            return torch.cat(self._hilo(w.permute(0, 2, 3, 1).reshape(co, 9 * ci)), dim=1)
        return self._put("conv3_hilo", key, f)

    def conv3_stem_hilo(self, key, c_pad=32):
        """stem conv3x3 weight [Cout, Cin <= 10, 3, 3] -> fp16 [Cout, 9*c_pad] for input channels laid out [x_hi | x_lo | x_hi | 0]
        (mvd_unet_input_f16 with hilo = 1): per tap [W_hi | W_hi | W_lo | 0]"""
        def f():
            w = self.raw(key).float()
            co, ci = w.shape[0], w.shape[1]
            assert 3 * ci <= c_pad
            hi, lo = self._hilo(w.permute(0, 2, 3, 1))  # [Cout, ky, kx, Cin]
            return torch.cat([hi, hi, lo, hi.new_zeros(co, 3, 3, c_pad - 3 * ci)], dim=3).reshape(co, 9 * c_pad)
        return self._put("conv3_stem_hilo", (key, c_pad), f)

    def qkv(self, p):
        """to_q | to_k | to_v fused into one [3C, C] matrix (external/sd1/ldm/modules/attention.py:161-163)"""
        return self._put("qkv", p, lambda: self._pad_k(torch.cat([self.raw(p + ".to_q.weight"), self.raw(p + ".to_k.weight"),
                                                                  self.raw(p + ".to_v.weight")], 0).float()).half())

    # nn.LayerNorm folded into the consuming GEMM (include/mvd_b200.h ABI 13):
    #   LN(x) W^T + b = rstd (x W'^T - mean colsum(W')) + (b + W beta),   W' = W diag(gamma)
    def _ln_fold(self, tag, key, w_fn, b_fn, norm):
        """-> (W' fp16 [N, K8], colsum fp32 [N] of the ROUNDED W', bias' fp32 [N])"""
        wp = self._put(tag + "_w", (key, norm), lambda: self._pad_k(w_fn().float() * self.raw(norm + ".weight").float()[None, :]).half())
        cs = self._put(tag + "_cs", (key, norm), lambda: wp.float().sum(dim=1))

        def f_b():
            wb = w_fn().double() @ self.raw(norm + ".bias").double()
            b = b_fn()
            return (wb if b is None else wb + b.double()).float()
        bp = self._put(tag + "_b", (key, norm), f_b)
        return wp, cs, bp

    def qkv_ln(self, p, norm):
        """fused to_q | to_k | to_v behind nn.LayerNorm `norm` (attention.py:211,220; mvd attention.py:35,52)"""
        w = lambda: torch.cat([self.raw(p + ".to_q.weight"), self.raw(p + ".to_k.weight"), self.raw(p + ".to_v.weight")], 0)
        return self._ln_fold("qkvln", p, w, lambda: None, norm)

    def geglu_ln(self, p, norm, tile_n=128):
        """GEGLU.proj behind nn.LayerNorm `norm`, rows interleaved per output tile as in geglu()"""
        inner = self.raw(p + ".weight").shape[0] // 2
        perm = self._put("geglu_perm", (inner, tile_n), lambda: self.ops.geglu_permutation(inner, tile_n))
        w = lambda: self.raw(p + ".weight")[perm.cpu()]
        b = lambda: self.raw(p + ".bias")[perm.cpu()]
        return self._ln_fold("gegluln", (p, tile_n), w, b, norm)

    def kv(self, p):
        return self._put("kv", p, lambda: self._pad_k(torch.cat([self.raw(p + ".to_k.weight"),
                                                                 self.raw(p + ".to_v.weight")], 0).float()).half())

    def geglu(self, p, tile_n=128):
        """GEGLU.proj [2*inner, C] with value/gate rows interleaved per output tile (include/mvd_b200.h, MVD_ACT_GEGLU)"""
        w = self.raw(p + ".weight")
        inner = w.shape[0] // 2
        perm = self._put("geglu_perm", (inner, tile_n), lambda: self.ops.geglu_permutation(inner, tile_n))
        wp = self._put("geglu_w", p, lambda: self._pad_k(self.raw(p + ".weight").float()[perm.cpu()]).half())
        bp = self._put("geglu_b", p, lambda: self.raw(p + ".bias").float()[perm.cpu()])
        return wp, bp


# ------------------------------------------------------------------------------------------------ buffers
class Arena:
    """Scratch-buffer pool.  Calls of a Program execute in order on one stream, so a buffer may be handed out
    again as soon as the emitter that produced its last reader has been appended."""

    def __init__(self, ops):
        self.ops = ops
        self.free_list = []   # base uint8 tensors
        self.base_of = {}     # data_ptr -> base
        self.total_bytes = 0

    def alloc(self, shape, dtype):
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = _round_up(max(n, 1) * torch.empty((), dtype=dtype).element_size(), 1024)
        best = None
        for i, b in enumerate(self.free_list):
            if b.numel() >= nbytes and (best is None or b.numel() < self.free_list[best].numel()):
                best = i
        if best is not None and self.free_list[best].numel() <= 2 * nbytes:
            base = self.free_list.pop(best)
        else:
            base = self.ops.empty((nbytes,), torch.uint8)
            self.total_bytes += nbytes
        t = base[: n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(*shape)
        self.base_of[t.data_ptr()] = base
        return t

    def free(self, *tensors):
        for t in tensors:
            if t is None:
                continue
            base = self.base_of.pop(t.data_ptr(), None)
            if base is not None:
                self.free_list.append(base)


class HostCall:
    """A host-side step of a Program that is not a C-ABI kernel call (e.g. the NCCL all-gather of the view-sharded mode,
    issued through torch.distributed on the current stream — capturable into the step's CUDA graph like the kernels)."""

    def __init__(self, name, fn):
        self.name, self.fn = name, fn
        self.meta = {"kernel": name}

    def __call__(self, stream):
        self.fn(stream)


_NVTX = bool(os.environ.get("MVD_NVTX"))


class StageMark:
    """Stage boundary inside a Program.  With MVD_NVTX=1 in the environment every stage of an eager (non-graph) replay is an NVTX
    range (`ncu --nvtx --nvtx-include "unet.middle_block/"`, or a timeline tool); otherwise a no-op.  Not a kernel launch."""

    _open = False

    def __init__(self, name):
        self.name = "stage:" + name
        self.stage = name
        self.meta = {"kernel": "stage_mark"}

    def __call__(self, stream):
        if not _NVTX:
            return
        if StageMark._open:
            torch.cuda.nvtx.range_pop()
        StageMark._open = self.stage != "end"
        if StageMark._open:
            torch.cuda.nvtx.range_push(self.stage)


class Program:
    """A fixed sequence of bound kernel calls."""

    def __init__(self):
        self.calls = []

    def mark(self, stage):
        """start a named stage (NVTX range when MVD_NVTX=1); only recorded when the instrumentation is on, so that the default
        program is the bare launch list"""
        if _NVTX:
            self.calls.append(StageMark(stage))

    def append(self, call):
        self.calls.append(call)

    def extend(self, other):
        self.calls.extend(other.calls)

    def run(self, stream):
        for c in self.calls:
            c(stream)

    def __len__(self):
        return len(self.calls)


# ------------------------------------------------------------------------------------------------ emitters
class Builder:
    """Emits kernel-call sequences for the reference modules into a Program."""

    def __init__(self, ops, weights, program=None, arena=None):
        self.ops = ops
        self.W = weights
        self.prog = program if program is not None else Program()
        self.arena = arena if arena is not None else Arena(ops)
        self._qkv_bufs = {}
        self._emb_vectors = {}
        self._stats = None
        self._ws = None
        self.heads = 8
        # fp16 operands of the 1x1 skip convolutions written by their producers' epilogues (MVD_NO_FUSE_CAT=1: cast / concat passes)
        self.fuse_cat = not os.environ.get("MVD_NO_FUSE_CAT")
        # split-precision (hi/lo fp16) operands where operand rounding lands on the residual trunk:
        # level 1 = stem + head convolutions, 2 = also the ResBlock 1x1 skip convolutions; MVD_HILO=0 turns it off (A/B)
        self.hilo = int(os.environ.get("MVD_HILO", "2"))
        # data movement folded into TMA addressing: strided implicit-GEMM Downsample; MVD_NO_FOLD_K8=1 = im2col (A/B)
        self.fold_k8 = not os.environ.get("MVD_NO_FOLD_K8")
        # nn.LayerNorm (norm1 / norm3 of the transformer blocks) without a pass of its own: the GEMM that produces the rows leaves their
        # fp16 copy and per-chunk (sum, sum of squares), the QKV / GEGLU GEMM folds the normalisation (ABI 13); MVD_NO_LN_FOLD=1: ln_kernel
        self.ln_fold = not os.environ.get("MVD_NO_LN_FOLD")
        # nearest x2 upsample folded into the following convolution (ABI 14): four 2 x 2 phase convolutions of the source image,
        # no upsampled tensor, 16 C instead of 36 C multiply-adds per output value; MVD_NO_FOLD_UP=1 = upsample pass + conv3x3 (A/B)
        self.fold_up = not os.environ.get("MVD_NO_FOLD_UP")
        # ... where it pays: up to 4096 rows.  Above that (the 32^2 level of an 8-view step, M = 16384) the consumers' per-unit epilogue work
        # outweighs a LayerNorm pass that streams at L2 speed (measured, profiles/r02_bench_v6_ab_lnfold_rows.md); MVD_LN_FOLD_MAX_ROWS overrides
        self.ln_fold_max_rows = int(os.environ.get("MVD_LN_FOLD_MAX_ROWS", "4096"))

    # -- buffers
    def t16(self, *shape):
        return self.arena.alloc(shape, torch.float16)

    def t32(self, *shape):
        return self.arena.alloc(shape, torch.float32)

    def free(self, *ts):
        self.arena.free(*ts)

    def stats_ws(self, n_img):
        if self._stats is None or self._stats.numel() < n_img * 64:
            self._stats = self.ops.zeros((max(n_img, 64) * 64,), torch.float64)
        return self._stats

    def qkv_buffers(self, n_img, seq, dpad):
        """q, k [n_img*heads, seq, dpad] and v^T [n_img*heads, dpad, seq]; head-dim padding stays zero because
        the QKV GEMM epilogue only ever writes the first dhead columns."""
        key = (n_img, seq, dpad)
        if key not in self._qkv_bufs:
            n = n_img * self.heads * seq * dpad
            self._qkv_bufs[key] = tuple(self.ops.zeros((n,), torch.float16) for _ in range(3))
        return self._qkv_bufs[key]

    # -- primitive emitters
    def splitk_ws(self):
        """Zero-initialised split-K workspace shared by every GEMM of the program (calls run in stream order; the tile
        semaphores at its head reset themselves)."""
        if self._ws is None:
            self._ws = self.ops.zeros((SPLITK_WS_BYTES,), torch.uint8)
        return self._ws

    def gemm(self, A, Wt, out, M, N, K, *, allow_split=False, **kw):
        # split_k = 0 lets the library cut K when the tile grid cannot fill the machine (small-M, weight-bound layers);
        # a measured (tile_n, split_k, cta_pair) choice from gemm_tuning.json overrides the library's heuristics.
        act = kw.get("act", ACT_NONE)
        can_split = allow_split and act != ACT_GEGLU and kw.get("qkv") is None and kw.get("ln_stats_out") is None
        sig = gemm_signature(kw.get("conv") is not None, M, N, K * (3 if kw.get("hilo") else 1),  # hi/lo operands: three passes over K
                             "qkv" if kw.get("qkv") is not None else str(out.dtype).split(".")[-1], kw.get("residual") is not None, act,
                             "+st" if kw.get("ln_stats_out") is not None else ("+ln" if kw.get("ln") is not None else ("+up" if kw.get("conv_up2") else "")))
        tuned = gemm_tuning().get(sig)
        if tuned is None and sig[-3:] in ("+st", "+ln"):  # not measured in this form yet: the choice of the plain form
            tuned = gemm_tuning().get(sig[:-3])
        if tuned is not None and "tile_n" not in kw:
            tn, sk, pr = tuned
            if act == ACT_GEGLU:
                tn = GEGLU_TILE
            if sk > 1 and not can_split:
                sk = 1
            self.prog.append(self.ops.gemm(A, Wt, out, M, N, K, split_k=sk, tile_n=tn, cta_pair=pr,
                                           ws=self.splitk_ws() if sk != 1 else None, **kw))
        elif can_split:
            self.prog.append(self.ops.gemm(A, Wt, out, M, N, K, split_k=0, ws=self.splitk_ws(), **kw))
        else:
            self.prog.append(self.ops.gemm(A, Wt, out, M, N, K, split_k=1, **kw))

    def groupnorm(self, x, p, n_img, hw, C, eps, silu):
        """GroupNorm32(32, C) (+SiLU) -> fp16 operand.  util.py:200-217 / attention.py:76-77"""
        y = self.t16(n_img * hw, C)
        self.prog.append(self.ops.groupnorm(x, self.W.f32(p + ".weight"), self.W.f32(p + ".bias"), y, None, n_img, hw, C, eps, silu))
        return y

    def layernorm(self, x, p, rows, C):
        y = self.t16(rows, C)
        self.prog.append(self.ops.layernorm(x, self.W.f32(p + ".weight"), self.W.f32(p + ".bias"), y, rows, C, 1e-5))
        return y

    def conv3x3(self, a16, wkey, n_img, H, W_, Cin, Cout, *, bias, residual=None, c_pad=None, ldc=None, out=None,
                rowbias=None, rows_per_group=1, out16=None, hilo=False, w=None):
        """conv_nd(2, Cin, Cout, 3, padding=1) as implicit GEMM (openaimodel.py:107,204,230; unet.py:323,499).
        hilo: a16 holds 2*Cin channels [hi | lo] and the weights are packed [W_hi | W_lo] (split-precision operands)."""
        M = n_img * H * W_
        cp = c_pad if c_pad is not None else Cin
        if out is None:
            out = self.t32(M, ldc if ldc is not None else Cout)
        if w is None:
            w = self.W.conv3_hilo(wkey) if hilo else self.W.conv3(wkey, c_pad)
        self.gemm(a16, w, out, M, Cout, 9 * cp, allow_split=True, conv=(n_img, H, W_, cp), bias=bias, residual=residual,
                  rowbias=rowbias, rows_per_group=rows_per_group, hilo=hilo,
                  ldr=(residual.shape[-1] if residual is not None else 0), ldc=(ldc if ldc is not None else Cout), **self._o16(out16))
        return out

    @staticmethod
    def _o16(out16):
        """out16 = (fp16 tensor or column window, row pitch[, column offset of the lo half]) -> gemm keywords: the GEMM's epilogue
        also writes its result there (and, for a [hi | lo] buffer, the fp16 rounding residual)"""
        if out16 is None:
            return {}
        return {"out16": out16[0], "ld16": out16[1], "out16_lo": out16[2] if len(out16) > 2 else 0}

    def cast16(self, x, rows, C):
        y = self.t16(rows, C)
        self.prog.append(self.ops.cast(x, y, rows * C))
        return y

    def emb_vectors(self, emb, emb_dim, blocks):
        """emb_layers (SiLU -> Linear) of every ResBlock in `blocks` [(prefix, C_out)] on the shared timestep embedding,
        as ONE grouped launch (openaimodel.py:218-224,266-270); the conv bias in_layers.2.bias is folded in."""
        total = sum(c for _, c in blocks)
        buf = self.ops.empty((total,), torch.float32)
        jobs, off = [], 0
        for p, c in blocks:
            y = buf[off:off + c]
            off += c
            jobs.append((self.W.lin(p + ".emb_layers.1.weight"), self.W.f32_sum(p + ".emb_layers.1.bias", p + ".in_layers.2.bias"), y))
            self._emb_vectors[p] = y.view(1, c)
        self.prog.append(self.ops.gemv_grouped(emb, emb_dim, jobs, silu_in=True))

    # -- ResBlock
    def resblock(self, x, p, n_img, H, Cin, Cout, emb, emb_dim, cat=None, x16=None, out16=None):
        """ResBlock._forward (openaimodel.py:255-275): GN-SiLU-conv, + Linear(SiLU(emb)), GN-SiLU-conv, + skip.
        The timestep term is per-channel only (one shared t), so it rides in the first conv's bias.
        cat = (h, skip): the block input is torch.cat([h, skip], dim=1) (mvdfusion/unet.py:550); it is consumed by a
        two-source GroupNorm and written once as the fp16 operand of the 1x1 skip convolution, never in fp32.
        x16 = (fp16 tensor / column window, row pitch): the block input as fp16, already written by its producers' epilogues
        (out16) — the skip convolution reads it instead of a cast / concat pass.  out16: where to leave this block's output."""
        hw, M = H * H, n_img * H * H
        if cat is not None:
            assert x is None and self.W.has(p + ".skip_connection.weight")
            c1, c2 = cat[0].shape[-1], cat[1].shape[-1]
            a = self.t16(M, Cin)
            self.prog.append(self.ops.groupnorm2(cat[0], c1, cat[1], c2, self.W.f32(p + ".in_layers.0.weight"),
                                                 self.W.f32(p + ".in_layers.0.bias"), a, n_img, hw, 1e-5, True))
        else:
            a = self.groupnorm(x, p + ".in_layers.0", n_img, hw, Cin, 1e-5, True)
        ne = emb.shape[0]  # 1 on the path (shared t); n_img when a caller passes per-image embeddings
        pre = self._emb_vectors.get(p) if ne == 1 else None
        if pre is not None:  # computed with every other ResBlock's vector by one grouped launch (emb_vectors)
            eb = pre
        else:
            eb = self.t32(ne, Cout)
            self.prog.append(self.ops.gemv(emb, self.W.lin(p + ".emb_layers.1.weight"),
                                           self.W.f32_sum(p + ".emb_layers.1.bias", p + ".in_layers.2.bias"), eb, ne, Cout,
                                           emb_dim, silu_in=True))
        if ne == 1:
            h = self.conv3x3(a, p + ".in_layers.2.weight", n_img, H, H, Cin, Cout, bias=eb)
        else:
            h = self.conv3x3(a, p + ".in_layers.2.weight", n_img, H, H, Cin, Cout, bias=None, rowbias=eb, rows_per_group=hw)
        self.free(a)
        b = self.groupnorm(h, p + ".out_layers.0", n_img, hw, Cout, 1e-5, True)
        self.free(h, eb)
        res, s = x, None
        if self.W.has(p + ".skip_connection.weight"):
            s = self.t32(M, Cout)
            if x16 is not None and len(x16) > 2 and x16[2]:  # [hi | lo] operand: split-precision skip convolution
                self.gemm(x16[0], self.W.lin_hilo(p + ".skip_connection.weight"), s, M, Cout, Cin, allow_split=True,
                          bias=self.W.f32(p + ".skip_connection.bias"), lda=x16[1], hilo=True, a_lo_off=x16[2])
            elif x16 is not None:
                self.gemm(x16[0], self.W.lin(p + ".skip_connection.weight"), s, M, Cout, Cin, allow_split=True,
                          bias=self.W.f32(p + ".skip_connection.bias"), lda=x16[1])
            else:
                if cat is not None:
                    xc = self.t16(M, Cin)
                    self.prog.append(self.ops.concat16(cat[0], cat[1], xc, M, cat[0].shape[-1], cat[1].shape[-1]))
                else:
                    xc = self.cast16(x, M, Cin)
                self.gemm(xc, self.W.lin(p + ".skip_connection.weight"), s, M, Cout, Cin, allow_split=True,
                          bias=self.W.f32(p + ".skip_connection.bias"))
                self.free(xc)
            res = s
        out = self.conv3x3(b, p + ".out_layers.3.weight", n_img, H, H, Cout, Cout, bias=self.W.f32(p + ".out_layers.3.bias"),
                           residual=res, out16=out16)
        self.free(b, s)
        return out

    def downsample(self, x, p, n_img, H, C, out16=None, x16=None):
        """Downsample.op: conv3x3 stride 2 (openaimodel.py:151).  With the input available as fp16 (x16 = (tensor / column window,
        pixel pitch[, lo offset]) written by its producer's epilogue) it is a strided implicit GEMM: the taps come straight from the
        full-resolution image through a TMA box with element strides 2 — no im2col pass, no [M, 9C] matrix.  Otherwise fp16 im2col + GEMM."""
        Mo = n_img * (H // 2) * (H // 2)
        if x16 is not None and self.fold_k8:
            out = self.t32(Mo, C)
            self.gemm(x16[0], self.W.conv3(p + ".op.weight"), out, Mo, C, 9 * C, allow_split=True, conv=(n_img, H // 2, H // 2, C),
                      conv_stride=2, lda=x16[1], bias=self.W.f32(p + ".op.bias"), **self._o16(out16))
            return out
        col = self.t16(Mo, 9 * C)
        self.prog.append(self.ops.im2col_s2(x, col, n_img, H, H, C))
        out = self.t32(Mo, C)
        self.gemm(col, self.W.conv3(p + ".op.weight"), out, Mo, C, 9 * C, allow_split=True, bias=self.W.f32(p + ".op.bias"),
                  **self._o16(out16))
        self.free(col)
        return out

    def up_ok(self, H, C):
        return self.fold_up and H & (H - 1) == 0 and C % 32 == 0

    def upsample(self, x, p, n_img, H, C, out16=None, x16=None):
        """Upsample: nearest x2 then conv3x3 (openaimodel.py:107-119).  With the input available as fp16 (x16 = (tensor, pixel pitch) written
        by its producer's epilogue) the upsample is folded into the convolution: four 2 x 2 phase convolutions of the SOURCE image in one
        launch (conv_up2) — no upsampled tensor, 2.25 x fewer multiply-adds."""
        if x16 is not None and self.up_ok(H, C):
            Mo = n_img * 4 * H * H
            out = self.t32(Mo, C)
            self.gemm(x16[0], self.W.conv3_up2(p + ".conv.weight"), out, n_img * H * H, 4 * C, 4 * C, allow_split=True, conv=(n_img, H, H, C),
                      conv_up2=True, lda=x16[1], bias=self.W.f32_tiled(p + ".conv.bias", 4), ldc=C, **self._o16(out16))
            return out
        u = self.t16(n_img * 4 * H * H, C)
        self.prog.append(self.ops.upsample2x(x, u, n_img, H, H, C))
        out = self.conv3x3(u, p + ".conv.weight", n_img, 2 * H, 2 * H, C, C, bias=self.W.f32(p + ".conv.bias"), out16=out16)
        self.free(u)
        return out

    # -- attention blocks
    def ln_ok(self, M, C):
        """can the LayerNorm over rows [M, C] ride in its neighbours' epilogues?  (ABI 13: whole 32-column chunks, the TMA epilogue)"""
        return self.ln_fold and M <= self.ln_fold_max_rows and C % 32 == 0 and C <= 1536 and (C // self.heads) % 8 == 0

    def ln_pair(self, M, C):
        """-> keywords for the GEMM that produces rows a folded LayerNorm will read: their fp16 copy + the per-chunk statistics"""
        h16, st = self.t16(M, C), self.t32(C // 32, M, 2)
        return (h16, st), {"out16": h16, "ld16": C, "ln_stats_out": st}

    def self_attention(self, h, p, norm, n_img, seq, C, rowbias=None, h_ln=None, want_ln=False):
        """x = attn1(norm1(x)) + x  [+ per-image vector]: LayerNorm -> fused QKV GEMM (heads scattered) ->
        flash attention -> to_out GEMM with bias + residual.  attention.py:170-193,220; mvd attention.py:52
        h_ln = (fp16 copy of h, chunk statistics) left by h's producer: norm1 is folded into the QKV GEMM (no LayerNorm pass);
        want_ln: also return such a pair for the result (for the next folded LayerNorm)."""
        M = n_img * seq
        d = C // self.heads
        dpad = _round_up(d, 64)
        q, k, vt = self.qkv_buffers(n_img, seq, dpad)
        qkv = dict(out_k=k, out_vt=vt, heads=self.heads, dhead=d, dpad=dpad, seq=seq)
        if h_ln is not None:
            w, cs, bq = self.W.qkv_ln(p, norm)
            self.gemm(h_ln[0], w, q, M, 3 * C, C, qkv=qkv, bias=bq, ln=(h_ln[1], cs, 1e-5))
            self.free(h_ln[1])
            ao = h_ln[0]  # reuse: same shape / dtype, the QKV GEMM was its last reader
        else:
            ao = self.layernorm(h, norm, M, C)
            self.gemm(ao, self.W.qkv(p), q, M, 3 * C, C, qkv=qkv)
        self.prog.append(self.ops.attn_self(q, k, vt, ao, n_img, self.heads, seq, d, dpad, C))
        h2 = self.t32(M, C)
        pair, kw = self.ln_pair(M, C) if want_ln else (None, {})
        self.gemm(ao, self.W.lin(p + ".to_out.0.weight"), h2, M, C, C, allow_split=True, bias=self.W.f32(p + ".to_out.0.bias"),
                  rowbias=rowbias, rows_per_group=seq, residual=h, ldr=C, **kw)
        self.free(ao, h)
        return (h2, pair) if want_ln else h2

    def feed_forward(self, h, p, norm, M, C, out16=False, h_ln=None):
        """x = ff(norm3(x)) + x with the GEGLU fused into the first GEMM's epilogue.  attention.py:37-64,222
        out16: the block output only feeds proj_out, so it is written once, as that GEMM's fp16 operand.
        h_ln = (fp16 copy of h, chunk statistics) left by h's producer: norm3 is folded into the GEGLU GEMM."""
        inner = 4 * C
        g = self.t16(M, inner)
        if h_ln is not None:
            wg, cs, bg = self.W.geglu_ln(p + ".net.0.proj", norm, GEGLU_TILE)
            self.gemm(h_ln[0], wg, g, M, 2 * inner, C, bias=bg, act=ACT_GEGLU, tile_n=GEGLU_TILE, ldc=inner, ln=(h_ln[1], cs, 1e-5))
            self.free(*h_ln)
        else:
            ln = self.layernorm(h, norm, M, C)
            wg, bg = self.W.geglu(p + ".net.0.proj", GEGLU_TILE)
            self.gemm(ln, wg, g, M, 2 * inner, C, bias=bg, act=ACT_GEGLU, tile_n=GEGLU_TILE, ldc=inner)
            self.free(ln)
        h2 = self.t16(M, C) if out16 else self.t32(M, C)
        self.gemm(g, self.W.lin(p + ".net.2.weight"), h2, M, C, inner, allow_split=True, bias=self.W.f32(p + ".net.2.bias"),
                  residual=h, ldr=C)
        self.free(g, h)
        return h2

    def spatial_transformer(self, x, p, n_img, H, C, clipvec, out16=None):
        """SpatialTransformer.forward (attention.py:268-287) with BasicTransformerBlock (:219-223).
        attn2 sees ONE CLIP token per view, so softmax == 1 and its output is the per-view vector
        to_out(to_v(ctx)) (`clipvec`, [n_img, C]); it is added in the attn1 output GEMM's epilogue."""
        hw, M = H * H, n_img * H * H
        a = self.groupnorm(x, p + ".norm", n_img, hw, C, 1e-6, False)
        h = self.t32(M, C)
        fold = self.ln_ok(M, C)
        pair, kw = self.ln_pair(M, C) if fold else (None, {})
        self.gemm(a, self.W.lin(p + ".proj_in.weight"), h, M, C, C, allow_split=True, bias=self.W.f32(p + ".proj_in.bias"), **kw)
        self.free(a)
        tb = p + ".transformer_blocks.0"
        if fold:
            h, pair = self.self_attention(h, tb + ".attn1", tb + ".norm1", n_img, hw, C, rowbias=clipvec, h_ln=pair, want_ln=True)
        else:
            h = self.self_attention(h, tb + ".attn1", tb + ".norm1", n_img, hw, C, rowbias=clipvec)
        h16 = self.feed_forward(h, tb + ".ff", tb + ".norm3", M, C, out16=True, h_ln=pair)
        out = self.t32(M, C)
        self.gemm(h16, self.W.lin(p + ".proj_out.weight"), out, M, C, C, allow_split=True, bias=self.W.f32(p + ".proj_out.bias"),
                  residual=x, ldr=C, **self._o16(out16))
        self.free(h16)
        return out

    def clip_vector(self, ctx, p, n_img, C, out=None):
        """Per-view output of the one-token CLIP cross-attention: to_out(to_v(ctx)) + bias  (attention.py:221)."""
        tb = p + ".transformer_blocks.0.attn2"
        v = self.t32(n_img, C)
        self.prog.append(self.ops.gemv(ctx, self.W.lin(tb + ".to_v.weight"), None, v, n_img, C, CTX_DIM))
        if out is None:
            out = self.ops.empty((n_img, C), torch.float32)
        self.prog.append(self.ops.gemv(v, self.W.lin(tb + ".to_out.0.weight"), self.W.f32(tb + ".to_out.0.bias"), out, n_img,
                                       C, C))
        self.free(v)
        return out

    def view_cross_attention(self, h, p, norm, ctx16, M, D, C, want_ln=False):
        """DualAttnetionBlock.attn2 (mvd attention.py:56-62): every pixel is one query against its D frustum keys.
        want_ln: also return (fp16 copy, chunk statistics) of the result for a folded LayerNorm behind it."""
        d = C // self.heads
        if D == 1:  # softmax over a single key == 1: out = to_out(to_v(ctx))
            v = self.t16(M, C)
            self.gemm(ctx16, self.W.lin(p + ".to_v.weight"), v, M, C, CTX_DIM)
            o = v
        else:
            ln = self.layernorm(h, norm, M, C)
            q = self.t16(M, C)
            self.gemm(ln, self.W.lin(p + ".to_q.weight"), q, M, C, C)
            self.free(ln)
            kv = self.t16(M * D, 2 * C)
            self.gemm(ctx16, self.W.kv(p), kv, M * D, 2 * C, CTX_DIM)
            o = self.t16(M, C)
            self.prog.append(self.ops.pixel_cross_attn(q, kv, o, M, D, self.heads, d))
            self.free(q, kv)
        h2 = self.t32(M, C)
        pair, kw = self.ln_pair(M, C) if want_ln else (None, {})
        self.gemm(o, self.W.lin(p + ".to_out.0.weight"), h2, M, C, C, allow_split=True, bias=self.W.f32(p + ".to_out.0.bias"),
                  residual=h, ldr=C, **kw)
        self.free(o, h)
        return (h2, pair) if want_ln else h2

    def view_aligned_transformer(self, x, p, n_img, H, C, ctx16, D, out16=None):
        """ViewAlignedFeatureTransformer.forward (mvd attention.py:119-145) + DualAttnetionBlock (:43-66)."""
        hw, M = H * H, n_img * H * H
        a = self.groupnorm(x, p + ".aligned_attn_norm", n_img, hw, C, 1e-6, False)
        h = self.t32(M, C)
        fold = self.ln_ok(M, C)
        pair, kw = self.ln_pair(M, C) if fold else (None, {})
        self.gemm(a, self.W.lin(p + ".aligned_attn_proj_in.weight"), h, M, C, C, allow_split=True,
                  bias=self.W.f32(p + ".aligned_attn_proj_in.bias"), **kw)
        self.free(a)
        tb = p + ".aligned_attn_transformer_blocks.0"
        h = self.self_attention(h, tb + ".attn1", tb + ".norm1", n_img, hw, C, h_ln=pair)
        if fold:
            h, pair = self.view_cross_attention(h, tb + ".attn2", tb + ".norm2", ctx16, M, D, C, want_ln=True)
        else:
            h = self.view_cross_attention(h, tb + ".attn2", tb + ".norm2", ctx16, M, D, C)
        h16 = self.feed_forward(h, tb + ".ff", tb + ".norm3", M, C, out16=True, h_ln=pair)
        out = self.t32(M, C)
        self.gemm(h16, self.W.lin(p + ".aligned_attn_proj_out.weight"), out, M, C, C, allow_split=True,
                  bias=self.W.f32(p + ".aligned_attn_proj_out.bias"), residual=x, ldr=C, **self._o16(out16))
        self.free(h16)
        return out

    # -- timestep MLPs
    def time_mlp(self, t_dev, freqs, dim, p0, p2, hidden, out_dim):
        """timestep_embedding -> Linear -> SiLU -> Linear on a single t (unet.py:537-538; viewfusion...py:276-279)"""
        te = self.t32(1, dim)
        self.prog.append(self.ops.timestep_embedding(t_dev, freqs, te, dim))
        e1 = self.t32(1, hidden)
        self.prog.append(self.ops.gemv(te, self.W.lin(p0 + ".weight"), self.W.f32(p0 + ".bias"), e1, 1, hidden, dim,
                                       silu_out=True))
        emb = self.ops.empty((1, out_dim), torch.float32)
        self.prog.append(self.ops.gemv(e1, self.W.lin(p2 + ".weight"), self.W.f32(p2 + ".bias"), emb, 1, out_dim, hidden))
        self.free(te, e1)
        return emb


def timestep_freqs(dim, device, max_period=10000):
    half = dim // 2
    return torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(device)


# ------------------------------------------------------------------------------------------------ UNet
class UNetSpec:
    """Topology of mvdfusion.unet.UNetModel (unet.py:320-500) as a flat op list derived from its hyper-parameters."""

    def __init__(self, model_channels, channel_mult, num_res_blocks, attention_resolutions, num_heads, image_size,
                 in_channels, out_channels):
        self.mc, self.mult, self.nres = model_channels, tuple(channel_mult), num_res_blocks
        self.attn_res = tuple(attention_resolutions)
        self.heads, self.image_size = num_heads, image_size
        self.in_channels, self.out_channels = in_channels, out_channels
        self.emb_dim = 4 * model_channels
        self.input_blocks, self.middle, self.output_blocks = [], [], []
        ch, ds = self.mc, 1
        chans = [ch]
        self.input_blocks.append([("stem", in_channels, ch)])
        for level, m in enumerate(self.mult):
            for _ in range(self.nres):
                layers = [("res", ch, m * self.mc)]
                ch = m * self.mc
                if ds in self.attn_res:
                    layers.append(("st", ch))
                self.input_blocks.append(layers)
                chans.append(ch)
            if level != len(self.mult) - 1:
                self.input_blocks.append([("down", ch)])
                chans.append(ch)
                ds *= 2
        self.middle = [("res", ch, ch), ("st", ch), ("vaft", ch), ("res", ch, ch)]
        for level, m in list(enumerate(self.mult))[::-1]:
            for i in range(self.nres + 1):
                ich = chans.pop()
                layers = [("res", ch + ich, self.mc * m)]
                ch = self.mc * m
                if ds in self.attn_res:
                    layers.append(("st", ch))
                    layers.append(("vaft", ch))
                if level and i == self.nres:
                    layers.append(("up", ch))
                    ds //= 2
                self.output_blocks.append(layers)
        self.final_ch = ch

    def st_layers(self):
        """(prefix, channels) of every SpatialTransformer in execution order."""
        out = []
        for name, blocks in (("input_blocks", self.input_blocks), ("output_blocks", self.output_blocks)):
            if name == "output_blocks":
                for j, l in enumerate(self.middle):
                    if l[0] == "st":
                        out.append((f"middle_block.{j}", l[1]))
            for i, layers in enumerate(blocks):
                for j, l in enumerate(layers):
                    if l[0] == "st":
                        out.append((f"{name}.{i}.{j}", l[1]))
        return out


def emit_unet(b, spec, x_in16, n_img, S, D, t_dev, freqs, clipvecs, pyramid16, c_in_pad=16, stem_hilo=False):
    """UNetModel.forward (unet.py:524-556).  x_in16: fp16 NHWC [n_img, S, S, c_in_pad]; clipvecs: dict prefix ->
    [n_img, C] fp32; pyramid16: list of fp16 [n_img*H_l*H_l*D, 768].  Returns the head output fp32 [n_img*S*S, 8].
    stem_hilo: x_in16 carries [x_hi | x_lo | x_hi | 0] channels (mvd_unet_input_f16 hilo) for a split-precision stem."""
    b.heads = spec.heads
    b.prog.mark("unet.time_embed")
    skip_hilo = b.hilo >= 2
    head_hilo = b.hilo >= 1 and spec.final_ch % 64 == 0
    emb = b.time_mlp(t_dev, freqs, spec.mc, "time_embed.0", "time_embed.2", spec.emb_dim, spec.emb_dim)
    res_blocks = []
    for name, blocks in (("input_blocks", spec.input_blocks), ("middle_block", [spec.middle]), ("output_blocks", spec.output_blocks)):
        for i, layers in enumerate(blocks):
            for j, l in enumerate(layers):
                if l[0] == "res":
                    res_blocks.append((f"{name}.{j}" if name == "middle_block" else f"{name}.{i}.{j}", l[2]))
    b.emb_vectors(emb, spec.emb_dim, res_blocks)
    H = S
    hs = []

    # The 1x1 skip convolution of output block i reads torch.cat([h, hs.pop()], dim=1) (unet.py:550) as an fp16 operand.
    # Both halves are written straight into that buffer by the epilogues of the GEMMs that produce h and the skip
    # (out16 windows), so no cast / concat pass runs; the channel-changing input ResBlocks read their input from the same
    # windows (row pitch c1 + c2).
    ch, hh, outs = spec.mc, S, []
    for layers in spec.input_blocks:
        for l in layers:
            if l[0] in ("stem", "res"):
                ch = l[2]
            elif l[0] == "down":
                hh //= 2
        outs.append((ch, hh))
    cat16 = []
    for i, layers in enumerate(spec.output_blocks):
        sch, sh = outs[len(outs) - 1 - i]
        assert sh == hh
        fused = b.fuse_cat and layers[0][0] == "res" and b.W.has(f"output_blocks.{i}.0.skip_connection.weight")
        wide = 2 if (skip_hilo and (ch + sch) % 64 == 0 and sch % 64 == 0) else 1  # [hi | lo]: the lo half starts at column ch + sch
        cat16.append((b.t16(n_img * hh * hh, wide * (ch + sch)), ch, sch, wide) if fused else None)
        for l in layers:
            if l[0] == "res":
                ch = l[2]
            elif l[0] == "up":
                hh *= 2

    def skip_window(k):
        """fp16 home of input block k's output: the skip half of its output block's concatenation buffer"""
        c = cat16[len(spec.input_blocks) - 1 - k]
        return None if c is None else (c[0][:, c[1]:], c[3] * (c[1] + c[2]), (c[1] + c[2]) if c[3] == 2 else 0)

    def head_window(i):
        """fp16 home of the tensor entering output block i (h half of its concatenation buffer)"""
        c = cat16[i] if i < len(cat16) else None
        return None if c is None else (c[0][:, :c[1]], c[3] * (c[1] + c[2]), (c[1] + c[2]) if c[3] == 2 else 0)

    def first_up16(layers, H):
        if not up_src_wanted(layers, 0, H):
            return None
        c_out = layers[0][2]
        up16[0] = (b.t16(n_img * H * H, c_out), c_out)
        return up16[0]

    up16 = [None]  # (fp16 copy of the tensor entering the next Upsample, pitch), written by the epilogue of the layer in front of it

    def up_src_wanted(layers, j, H):
        return j + 1 < len(layers) and layers[j + 1][0] == "up" and layers[j][0] in ("res", "st", "vaft") and b.up_ok(H, layers[j + 1][1])

    def run_layers(h, prefix, layers, H, start=0, h16=None, last16=None):
        for j, l in enumerate(layers):
            if j < start:
                continue
            p = f"{prefix}.{j}"
            kind = l[0]
            o16 = last16 if j == len(layers) - 1 else None
            if o16 is None and up_src_wanted(layers, j, H):  # the layer behind this one is an Upsample: leave it its input as fp16
                c_out = l[2] if kind == "res" else l[1]
                up16[0] = (b.t16(n_img * H * H, c_out), c_out)
                o16 = up16[0]
            if kind == "stem":
                new = b.conv3x3(h, p + ".weight", n_img, H, H, l[1], l[2], bias=b.W.f32(p + ".bias"), c_pad=c_in_pad, out16=o16,
                                w=b.W.conv3_stem_hilo(p + ".weight", c_in_pad) if stem_hilo else None)
            elif kind == "res":
                new = b.resblock(h, p, n_img, H, l[1], l[2], emb, spec.emb_dim, x16=h16 if j == 0 else None, out16=o16)
            elif kind == "st":
                new = b.spatial_transformer(h, p, n_img, H, l[1], clipvecs[p], out16=o16)
            elif kind == "vaft":
                level = {spec.image_size: 0, spec.image_size // 2: 1, spec.image_size // 4: 2, spec.image_size // 8: 3}[H]
                new = b.view_aligned_transformer(h, p, n_img, H, l[1], pyramid16[level], D, out16=o16)
            elif kind == "down":
                new = b.downsample(h, p, n_img, H, l[1], out16=o16, x16=h16 if j == 0 else None)
                H //= 2
            elif kind == "up":
                new = b.upsample(h, p, n_img, H, l[1], out16=o16, x16=up16[0])
                if up16[0] is not None:
                    b.free(up16[0][0])
                    up16[0] = None
                H *= 2
            else:
                raise ValueError(kind)
            if not any(h is s for s, _ in hs) and h is not x_in16:
                b.free(h)
            h = new
        return h, H

    h = x_in16
    for i, layers in enumerate(spec.input_blocks):
        b.prog.mark(f"unet.input_blocks.{i}")
        # a channel-changing ResBlock needs its input as fp16 (1x1 skip convolution): the previous block left it in its window
        needs16 = (layers[0][0] == "res" and layers[0][1] != layers[0][2]) or layers[0][0] == "down"
        prev16 = skip_window(i - 1) if i > 0 and needs16 else None
        h, H = run_layers(h, f"input_blocks.{i}", layers, H, h16=prev16, last16=skip_window(i))
        hs.append((h, layers))
    b.prog.mark("unet.middle_block")
    h, H = run_layers(h, "middle_block", spec.middle, H, last16=head_window(0))
    for i, layers in enumerate(spec.output_blocks):
        b.prog.mark(f"unet.output_blocks.{i}")
        skip, _ = hs.pop()
        c1, c2 = h.shape[-1], skip.shape[-1]
        rows = n_img * H * H
        p0 = f"output_blocks.{i}.0"
        if layers[0][0] == "res" and b.W.has(p0 + ".skip_connection.weight"):
            c16 = cat16[i]
            assert c16 is None or (c16[1], c16[2]) == (c1, c2)
            nxt = head_window(i + 1)
            new = b.resblock(None, p0, n_img, H, c1 + c2, layers[0][2], emb, spec.emb_dim, cat=(h, skip),
                             x16=None if c16 is None else (c16[0], c16[3] * (c1 + c2), (c1 + c2) if c16[3] == 2 else 0),
                             out16=nxt if len(layers) == 1 else first_up16(layers, H))
            b.free(h, skip)
            if c16 is not None:
                b.free(c16[0])
            h, H = run_layers(new, f"output_blocks.{i}", layers, H, start=1, last16=nxt)
        else:
            cat = b.t32(rows, c1 + c2)
            b.prog.append(b.ops.concat(h, skip, cat, rows, c1, c2))
            b.free(h, skip)
            h, H = run_layers(cat, f"output_blocks.{i}", layers, H, last16=head_window(i + 1))
    b.prog.mark("unet.out")
    if head_hilo:  # the head's operand rounding would reach the output unattenuated: split-precision operands
        a = b.t16(n_img * H * H, 2 * spec.final_ch)
        b.prog.append(b.ops.groupnorm_hilo(h, b.W.f32("out.0.weight"), b.W.f32("out.0.bias"), a, n_img, H * H, spec.final_ch, 1e-5, True))
    else:
        a = b.groupnorm(h, "out.0", n_img, H * H, spec.final_ch, 1e-5, True)
    b.free(h)
    head = b.ops.empty((n_img * H * H, 8), torch.float32)
    b.conv3x3(a, "out.2.weight", n_img, H, H, spec.final_ch, spec.out_channels, bias=b.W.f32("out.2.bias"), ldc=8, out=head,
              hilo=head_hilo)
    b.free(a)
    return head


# ------------------------------------------------------------------------------------------------ GridAttn
def emit_gridattn(b, *, noisy, input_latent, depth_override, depth_eps, scal, cams, mask, c_embed, n_views, S, D,
                  q_first, q_count, num_layers, num_heads, depth_scale, depth_shift, frustum_out, harm_freqs, ndc_grid):
    """GridAttn.forward + aggregate_features (mvdfusion/view_attn_efficient2.py:269-442) for query views
    [q_first, q_first+q_count) against all n_views views.  c_embed: fp32 [1, 256] (t_embed[:1]).
    frustum_out: fp16 [q_count*S*S*D, 768] (pyramid level 0 of the conditional images)."""
    hw = S * S
    V = n_views
    P = q_count * hw * D
    R = P * V
    W = b.W
    b.prog.mark("gridattn.geometry")
    feat = b.t16((n_views + 1) * hw, Z_CH)
    zdepth = b.t32(n_views * D * hw)
    b.prog.append(b.ops.gridattn_prep(noisy, input_latent, depth_override, depth_eps, scal, W.f32("z_embedder.0.weight"),
                                      W.f32("z_embedder.0.bias"), feat, zdepth, n_views, S, D, depth_scale, depth_shift))
    tokens = b.t16(R, TOKEN_LD)
    b.prog.append(b.ops.gridattn_tokens(feat, zdepth, cams, mask, harm_freqs, ndc_grid, tokens, n_views, S, D, q_first, q_count))
    b.free(feat, zdepth)
    b.prog.mark("gridattn.transformer")
    hid0 = W.raw("aggregation_transformer.layer_list.0.mlp.fc1.weight").shape[0]
    # One kernel for pre_layer_b + the DiT blocks + the view pooling (csrc/dit.cu) whenever the tile geometry allows it: the V rows of
    # a point must sit inside one 128-row tile.  MVD_NO_DIT_FUSION=1 keeps the unfused program (A/B measurements).
    fused = (V & (V - 1)) == 0 and V <= 16 and num_heads == 8 and Z_CH == 256 and hid0 == 512 and 1 <= num_layers <= 4 and \
        not os.environ.get("MVD_NO_DIT_FUSION")
    if fused:
        mods = [b.ops.empty((6 * Z_CH,), torch.float32) for _ in range(num_layers)]
        b.prog.append(b.ops.gemv_grouped(c_embed, Z_CH, [(W.lin(f"aggregation_transformer.layer_list.{i}.adaLN_modulation.1.weight"),
                                                          W.f32(f"aggregation_transformer.layer_list.{i}.adaLN_modulation.1.bias"), mods[i])
                                                         for i in range(num_layers)], silu_in=True))
        layers, jobs = [], []
        for i in range(num_layers):
            p = f"aggregation_transformer.layer_list.{i}"
            ch = [mods[i][j * Z_CH:(j + 1) * Z_CH] for j in range(6)]  # shift / scale / gate (msa), shift / scale / gate (mlp)
            wp, wf = W.lin(p + ".attn.proj.weight"), W.lin(p + ".mlp.fc2.weight")
            wp_g, wf_g = b.ops.empty(tuple(wp.shape), torch.float16), b.ops.empty(tuple(wf.shape), torch.float16)
            bp_g, bf_g = b.ops.empty((Z_CH,), torch.float32), b.ops.empty((Z_CH,), torch.float32)
            jobs += [(wp, ch[2], W.f32(p + ".attn.proj.bias"), wp_g, bp_g), (wf, ch[5], W.f32(p + ".mlp.fc2.bias"), wf_g, bf_g)]
            layers.append({"w_qkv": W.qkv_heads(p + ".attn.qkv.weight", num_heads, Z_CH // num_heads),
                           "b_qkv": W.f32_heads(p + ".attn.qkv.bias", num_heads, Z_CH // num_heads),
                           "w_proj": wp_g, "b_proj": bp_g, "w_fc1": W.lin(p + ".mlp.fc1.weight"), "b_fc1": W.f32(p + ".mlp.fc1.bias"),
                           "w_fc2": wf_g, "b_fc2": bf_g, "shift_msa": ch[0], "scale_msa": ch[1], "shift_mlp": ch[3], "scale_mlp": ch[4]})
        for j0 in range(0, len(jobs), 8):
            b.prog.append(b.ops.dit_fold_gates(jobs[j0:j0 + 8]))
        pooled = b.t16(P, Z_CH)
        b.prog.append(b.ops.gridattn_dit(tokens, TOKEN_LD, W.lin("pre_layer_b.0.weight", k_pad=TOKEN_LD), W.f32("pre_layer_b.0.bias"), layers,
                                         W.f32("aggregation_transformer.weight_layer.weight"), W.f32("aggregation_transformer.weight_layer.bias"),
                                         pooled, R, V, 1e-6))
        b.free(tokens)
        out_dim = W.raw("final_layer_b.weight").shape[0]
        b.gemm(pooled, W.lin("final_layer_b.weight"), frustum_out, P, out_dim, Z_CH, bias=W.f32("final_layer_b.bias"))
        b.free(pooled)
        return
    x = b.t32(R, Z_CH)
    b.gemm(tokens, W.lin("pre_layer_b.0.weight", k_pad=TOKEN_LD), x, R, Z_CH, TOKEN_LD, bias=W.f32("pre_layer_b.0.bias"),
           act=ACT_GELU)
    b.free(tokens)
    hd = Z_CH // num_heads
    # adaLN_modulation (SiLU -> Linear(256, 6*256)) of every DiT block reads the same t embedding: one grouped launch
    mods = [b.ops.empty((6 * Z_CH,), torch.float32) for _ in range(num_layers)]  # shift/scale/gate (msa), shift/scale/gate (mlp)
    b.prog.append(b.ops.gemv_grouped(c_embed, Z_CH, [(W.lin(f"aggregation_transformer.layer_list.{i}.adaLN_modulation.1.weight"),
                                                      W.f32(f"aggregation_transformer.layer_list.{i}.adaLN_modulation.1.bias"), mods[i])
                                                     for i in range(num_layers)], silu_in=True))
    for i in range(num_layers):
        p = f"aggregation_transformer.layer_list.{i}"
        ch = [mods[i][j * Z_CH:(j + 1) * Z_CH] for j in range(6)]
        a = b.t16(R, Z_CH)
        b.prog.append(b.ops.ln_modulate(x, ch[0], ch[1], a, R, Z_CH, 1e-6))
        qkv = b.t16(R, 3 * Z_CH)
        b.gemm(a, W.lin(p + ".attn.qkv.weight"), qkv, R, 3 * Z_CH, Z_CH, bias=W.f32(p + ".attn.qkv.bias"))
        b.prog.append(b.ops.view_attention(qkv, a, P, V, num_heads, hd))
        b.free(qkv)
        x2 = b.t32(R, Z_CH)
        b.gemm(a, W.lin(p + ".attn.proj.weight"), x2, R, Z_CH, Z_CH, bias=W.f32(p + ".attn.proj.bias"), colscale=ch[2],
               residual=x, ldr=Z_CH)
        b.free(x)
        b.prog.append(b.ops.ln_modulate(x2, ch[3], ch[4], a, R, Z_CH, 1e-6))
        hid = W.raw(p + ".mlp.fc1.weight").shape[0]
        f = b.t16(R, hid)
        b.gemm(a, W.lin(p + ".mlp.fc1.weight"), f, R, hid, Z_CH, bias=W.f32(p + ".mlp.fc1.bias"), act=ACT_GELU)
        b.free(a)
        x = b.t32(R, Z_CH)
        b.gemm(f, W.lin(p + ".mlp.fc2.weight"), x, R, Z_CH, hid, bias=W.f32(p + ".mlp.fc2.bias"), colscale=ch[5],
               residual=x2, ldr=Z_CH)
        b.free(f, x2)
    pooled = b.t16(P, Z_CH)
    b.prog.append(b.ops.view_pool(x, W.f32("aggregation_transformer.weight_layer.weight"),
                                  W.f32("aggregation_transformer.weight_layer.bias"), pooled, P, V, Z_CH))
    b.free(x)
    out_dim = W.raw("final_layer_b.weight").shape[0]
    b.gemm(pooled, W.lin("final_layer_b.weight"), frustum_out, P, out_dim, Z_CH, bias=W.f32("final_layer_b.bias"))
    b.free(pooled)


def emit_pyramid(b, pyramid16, n_cond, S, D, C=CTX_DIM):
    """get_volume_feats_pyramid (unet.py:198-209): 'area' pooling of level 0 by 2, 4, 8 (conditional images only)."""
    for l in range(1, len(pyramid16)):
        b.prog.append(b.ops.frustum_pool(pyramid16[0], pyramid16[l], n_cond, S, D, C, 1 << l))


# ------------------------------------------------------------------------------------------------ VAE decoder
def _vae_resnet(b, x, p, n, H, Cin, Cout):
    """ResnetBlock.forward with temb = None (external/sd1/ldm/modules/diffusionmodules/model.py:122-141):
    GN(eps 1e-6)-swish-conv3x3, GN-swish-conv3x3, + x (through the 1x1 nin_shortcut when the width changes)."""
    W, hw, M = b.W, H * H, n * H * H
    a = b.groupnorm(x, p + ".norm1", n, hw, Cin, 1e-6, True)
    h = b.conv3x3(a, p + ".conv1.weight", n, H, H, Cin, Cout, bias=W.f32(p + ".conv1.bias"))
    b.free(a)
    a = b.groupnorm(h, p + ".norm2", n, hw, Cout, 1e-6, True)
    b.free(h)
    res, s = x, None
    if W.has(p + ".nin_shortcut.weight"):
        x16 = b.cast16(x, M, Cin)
        s = b.t32(M, Cout)
        b.gemm(x16, W.lin(p + ".nin_shortcut.weight"), s, M, Cout, Cin, allow_split=True, bias=W.f32(p + ".nin_shortcut.bias"))
        b.free(x16)
        res = s
    out = b.conv3x3(a, p + ".conv2.weight", n, H, H, Cout, Cout, bias=W.f32(p + ".conv2.bias"), residual=res)
    b.free(a, s, x)
    return out


def _vae_attn(b, x, p, n, S, C):
    """AttnBlock.forward (model.py:176-202): one head of width C over the S*S pixels of each image.  q | k | v are one fused
    GEMM whose epilogue leaves q, k as [n, seq, C] and v transposed as [n, C, seq]; per image S = q k^T (fp32), row softmax
    with the c^-1/2 scale, O = P V; proj_out adds the residual."""
    W, seq = b.W, S * S
    M = n * seq
    a = b.groupnorm(x, p + ".norm", n, seq, C, 1e-6, False)
    wqkv = W._put("vae_qkv_w", p, lambda: W._pad_k(torch.cat([W.raw(p + ".q.weight"), W.raw(p + ".k.weight"), W.raw(p + ".v.weight")], 0)
                                                 .float().reshape(3 * C, C)).half())
    bqkv = W._put("vae_qkv_b", p, lambda: torch.cat([W.raw(p + ".q.bias"), W.raw(p + ".k.bias"), W.raw(p + ".v.bias")], 0).float())
    q, k, vt = b.t16(M, C), b.t16(M, C), b.t16(n * C, seq)
    b.gemm(a, wqkv, q, M, 3 * C, C, bias=bqkv, qkv=dict(out_k=k, out_vt=vt, heads=1, dhead=C, dpad=C, seq=seq))
    b.free(a)
    sc, pr, o = b.t32(seq, seq), b.t16(seq, seq), b.t16(M, C)
    for i in range(n):
        b.gemm(q[i * seq:(i + 1) * seq], k[i * seq:(i + 1) * seq], sc, seq, seq, C)
        b.prog.append(b.ops.softmax_rows(sc, pr, seq, seq, float(C) ** -0.5))
        b.gemm(pr, vt[i * C:(i + 1) * C], o[i * seq:(i + 1) * seq], seq, C, seq)
    b.free(q, k, vt, sc, pr)
    out = b.t32(M, C)
    b.gemm(o, W.lin(p + ".proj_out.weight"), out, M, C, C, allow_split=True, bias=W.f32(p + ".proj_out.bias"), residual=x, ldr=C)
    b.free(o, x)
    return out


def emit_vae_decoder(b, z_nchw, n, S, ch, ch_mult, num_res_blocks, z_channels=4, out_ch=3):
    """AutoencoderKL.decode (external/sd1/ldm/models/autoencoder.py:331-334) = post_quant_conv (1x1) -> Decoder.forward
    (model.py:541-577).  z_nchw: fp32 [n, z_channels, S*S].  Returns (fp32 rows [n*(S*2^(L-1))^2, 4] of which out_ch columns are
    written, output side).  The reference rounds the final normalised activation to fp16 before swish + conv_out
    (model.py:563-569); here GroupNorm + swish are one kernel that writes the fp16 conv operand, which differs from that
    by one fp16 rounding of an fp16-rounded value."""
    W, ops = b.W, b.ops
    M = n * S * S
    z16 = ops.zeros((M, 16), torch.float16)   # NHWC with the channel dim padded to 16 (padding stays zero)
    zq16 = ops.zeros((M, 16), torch.float16)
    b.prog.append(ops.nchw_to_nhwc16(z_nchw, z16, n, z_channels, S * S, 16))
    b.gemm(z16, W.lin("post_quant_conv.weight", k_pad=16), zq16, M, z_channels, 16, bias=W.f32("post_quant_conv.bias"), ldc=16)
    C = ch * ch_mult[-1]
    h = b.conv3x3(zq16, "decoder.conv_in.weight", n, S, S, z_channels, C, bias=W.f32("decoder.conv_in.bias"), c_pad=16)
    h = _vae_resnet(b, h, "decoder.mid.block_1", n, S, C, C)
    h = _vae_attn(b, h, "decoder.mid.attn_1", n, S, C)
    h = _vae_resnet(b, h, "decoder.mid.block_2", n, S, C, C)
    H = S
    for lvl in reversed(range(len(ch_mult))):
        Cout = ch * ch_mult[lvl]
        for j in range(num_res_blocks + 1):
            h = _vae_resnet(b, h, f"decoder.up.{lvl}.block.{j}", n, H, C, Cout)
            C = Cout
        if lvl != 0:
            u = b.upsample(h, f"decoder.up.{lvl}.upsample", n, H, C)
            b.free(h)
            h = u
            H *= 2
    a = b.groupnorm(h, "decoder.norm_out", n, H * H, C, 1e-6, True)
    b.free(h)
    out = ops.empty((n * H * H, 4), torch.float32)
    b.conv3x3(a, "decoder.conv_out.weight", n, H, H, C, out_ch, bias=W.f32("decoder.conv_out.bias"), ldc=4, out=out)
    b.free(a)
    return out, H


def emit_vae_encoder(b, x_nchw, n, R, ch, ch_mult, num_res_blocks, in_channels=3, z_channels=4, embed_dim=4):
    """AutoencoderKL.encode up to the moments (external/sd1/ldm/models/autoencoder.py:325-329) = Encoder.forward
    (model.py:440-460; double_z) -> quant_conv (1x1).  x_nchw: fp32 [n, in_channels, R*R] in [-1, 1].  Returns
    (fp32 rows [n*S*S, 8] = mean | logvar of the posterior, S)."""
    W, ops = b.W, b.ops
    x16 = ops.zeros((n * R * R, 16), torch.float16)  # NHWC, channel dim padded to 16 (padding stays zero)
    b.prog.append(ops.nchw_to_nhwc16(x_nchw, x16, n, in_channels, R * R, 16))
    h = b.conv3x3(x16, "encoder.conv_in.weight", n, R, R, in_channels, ch, bias=W.f32("encoder.conv_in.bias"), c_pad=16)
    C, H = ch, R
    for lvl in range(len(ch_mult)):
        Cout = ch * ch_mult[lvl]
        for j in range(num_res_blocks):
            h = _vae_resnet(b, h, f"encoder.down.{lvl}.block.{j}", n, H, C, Cout)
            C = Cout
        if lvl != len(ch_mult) - 1:
            # Downsample: F.pad(x, (0,1,0,1)) then conv3x3 stride 2 padding 0 (model.py:65-76) = im2col without low-side padding + GEMM
            p = f"encoder.down.{lvl}.downsample.conv"
            Mo = n * (H // 2) * (H // 2)
            col = b.t16(Mo, 9 * C)
            b.prog.append(ops.im2col_s2(h, col, n, H, H, C, pad_lo=0))
            b.free(h)
            h = b.t32(Mo, C)
            b.gemm(col, W.conv3(p + ".weight"), h, Mo, C, 9 * C, allow_split=True, bias=W.f32(p + ".bias"))
            b.free(col)
            H //= 2
    h = _vae_resnet(b, h, "encoder.mid.block_1", n, H, C, C)
    h = _vae_attn(b, h, "encoder.mid.attn_1", n, H, C)
    h = _vae_resnet(b, h, "encoder.mid.block_2", n, H, C, C)
    a = b.groupnorm(h, "encoder.norm_out", n, H * H, C, 1e-6, True)
    b.free(h)
    M, Nz = n * H * H, 2 * z_channels
    mom = b.t32(M, Nz)
    b.conv3x3(a, "encoder.conv_out.weight", n, H, H, C, Nz, bias=W.f32("encoder.conv_out.bias"), out=mom)
    b.free(a)
    m16 = b.cast16(mom, M, Nz)
    b.free(mom)
    out = ops.empty((M, 2 * embed_dim), torch.float32)
    b.gemm(m16, W.lin("quant_conv.weight"), out, M, 2 * embed_dim, Nz, bias=W.f32("quant_conv.bias"))
    b.free(m16)
    return out, H


# ------------------------------------------------------------------------------------------------ CLIP ViT image encoder
class ClipPlan:
    """OpenAI CLIP VisionTransformer.forward (clip/model.py; called through FrozenCLIPImageEmbedder, external/sd1/ldm/modules/
    encoders/modules.py:402-441) for B pre-processed images as one program.  Token layout: `seq` = T rounded up to a multiple of 16
    rows per image (ViT-L/14: T = 257 in 272 rows); rows [T, seq) are padding that the masked attention never reads as keys.
    W: PackedWeights over the `model.visual.` prefix.  Inputs: self.patches fp16 [B*(T-1), kp]; output: self.out fp32 [B, out_dim]."""

    def __init__(self, ops, W, B, width, layers, heads, T, kp, out_dim):
        self.ops = ops
        seq = _round_up(T, 16)
        d = width // heads
        dpad = _round_up(d, 64)
        b = Builder(ops, W)
        b.heads = heads
        self.patches = ops.zeros((B * (T - 1), kp), torch.float16)
        self.out = ops.empty((B, out_dim), torch.float32)
        M = B * seq
        # [class_embedding ; patches] + positional_embedding: the patch GEMM adds the positional rows as its residual and writes
        # straight into the token buffer (one launch per image: rows b*seq + 1 ..); row 0 of every image is a constant
        x = ops.zeros((M, width), torch.float32)
        pos = W.raw("positional_embedding").float()
        cls_row = (W.raw("class_embedding").float() + pos[0]).to(ops.device)
        pos_rest = pos[1:T].contiguous().to(ops.device)
        wconv = W._put("clip_conv1", kp, lambda: torch.nn.functional.pad(W.raw("conv1.weight").float().reshape(width, -1), (0, kp - 3 * W.raw("conv1.weight").shape[-1] ** 2)).half())
        x.view(B, seq, width)[:, 0].copy_(cls_row)
        self._keep = (cls_row, pos_rest)
        for i in range(B):
            b.gemm(self.patches[i * (T - 1):(i + 1) * (T - 1)], wconv, x[i * seq + 1:i * seq + T], T - 1, width, kp, residual=pos_rest, ldr=width)
        h = b.t32(M, width)
        b.prog.append(ops.layernorm_f32(x, W.f32("ln_pre.weight"), W.f32("ln_pre.bias"), h, M, width, 1e-5))
        self._x = x
        q, k, vt = b.qkv_buffers(B, seq, dpad)
        for l in range(layers):
            p = f"transformer.resblocks.{l}"
            ln = b.layernorm(h, p + ".ln_1", M, width)
            b.gemm(ln, W.lin(p + ".attn.in_proj_weight"), q, M, 3 * width, width, bias=W.f32(p + ".attn.in_proj_bias"),
                   qkv=dict(out_k=k, out_vt=vt, heads=heads, dhead=d, dpad=dpad, seq=seq))
            b.prog.append(ops.attn_self(q, k, vt, ln, B, heads, seq, d, dpad, width, seq_valid=T))
            h2 = b.t32(M, width)
            b.gemm(ln, W.lin(p + ".attn.out_proj.weight"), h2, M, width, width, bias=W.f32(p + ".attn.out_proj.bias"), residual=h, ldr=width)
            b.free(ln, h)
            ln = b.layernorm(h2, p + ".ln_2", M, width)
            # QuickGELU(u) = u sigmoid(1.702 u) = silu(1.702 u) / 1.702: c_fc is packed scaled by 1.702 (SiLU epilogue), c_proj by 1 / 1.702
            wfc = W._put("clip_fc", p, lambda p=p: (W.raw(p + ".mlp.c_fc.weight").float() * 1.702).half())
            bfc = W._put("clip_fc_b", p, lambda p=p: W.raw(p + ".mlp.c_fc.bias").float() * 1.702)
            wpr = W._put("clip_proj", p, lambda p=p: (W.raw(p + ".mlp.c_proj.weight").float() / 1.702).half())
            f = b.t16(M, 4 * width)
            b.gemm(ln, wfc, f, M, 4 * width, width, bias=bfc, act=ACT_SILU)
            b.free(ln)
            h = b.t32(M, width)
            b.gemm(f, wpr, h, M, width, 4 * width, bias=W.f32(p + ".mlp.c_proj.bias"), residual=h2, ldr=width)
            b.free(f, h2)
        # ln_post on the class tokens (row 0 of every image), then @ proj
        cls32 = b.t32(B, width)
        b.prog.append(ops.layernorm_f32(h, W.f32("ln_post.weight"), W.f32("ln_post.bias"), cls32, B, width, 1e-5, ldx=seq * width))
        wproj = W._put("clip_vproj", 0, lambda: W.raw("proj").float().t().contiguous().half())
        b.prog.append(ops.gemv(cls32, wproj, None, self.out, B, out_dim, width))
        self.prog = b.prog
