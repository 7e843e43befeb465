"""CPU: the SOURCE of the plain-SIMT kernels of the hot path — csrc/elementwise.cu (cast / concat / upsample / im2col / GEMV / grouped GEMV /
timestep embedding / UNet input assembly / CFG combine + DDIM update / layout converters / table gather) and csrc/gridattn.cu (GridAttn's
depth de-bias + z_embedder, the unproject -> reproject -> bilinear-gather -> harmonic-embedding token producer, the V x V view attention in
both variants, view pooling, frustum pooling, the D-key pixel cross-attention) and csrc/norm.cu (the cluster GroupNorm(+SiLU) in its
one- / two-source and split-precision forms — the blocks of a thread-block cluster run together on the shim and read each other's
shared memory, as over DSMEM — LayerNorm, adaLN modulate, fp32 LayerNorm with row pitches, row softmax) — compiled as C++ and executed on
host threads (tests/native/cpu_emul/cuda_on_cpu.h), through the product's own ops.NativeOps bindings, against the same emulations of the
documented semantics the B200 tests use — the very test bodies of tests/test_gpu_ops.py, with `nat` bound to the shim build.
(The three inline-PTX spots of gridattn.cu — fma.rn.f32.f16 and ex2.approx — have plain C++ stand-ins under MVD_CPU_EMULATION.)

What this proves without a GPU: indexing, tails, fp16 rounding (IEEE binary16 through _Float16), the grouped-GEMV job table, the DDIM /
CFG arithmetic, GroupNorm slice / cluster geometry and statistics exchange.  What it cannot prove: anything about tcgen05 / TMA
(gemm.cu, attention.cu, dit.cu are hardware-only), memory ordering of the real cluster barrier, or timing; the -m gpu suite remains the
proof on the B200."""
import os

import pytest

import test_gpu_ops as G
from common import build_cpu_shim, shim_ops


@pytest.fixture(scope="module")
def shim_lib(tmp_path_factory):
    return build_cpu_shim(["elementwise.cu", "gridattn.cu", "norm.cu"], tmp_path_factory.mktemp("cpu_shim"), "libmvd_simt_cpuemul.so")


@pytest.fixture
def nat(shim_lib, monkeypatch):
    return shim_ops(shim_lib, monkeypatch)


@pytest.fixture
def dbl():
    from ops_double import TorchOpsDouble
    return TorchOpsDouble()


def test_data_movement_kernels(nat, dbl):
    G.test_data_movement(nat, dbl)


def test_concat16_kernel(nat, dbl):
    t = {"a": G.rnd(70, 64), "b": G.rnd(70, 32, seed=1), "c": G.torch.zeros(70, 96, dtype=G.torch.float16)}
    G.run_both(nat, dbl, "concat16", t, ["c"], "a", "b", "c", 70, 64, 32, tol=1e-3)


def test_gemv_grouped_kernel(nat, dbl):
    G.test_gemv_grouped(nat, dbl)


def test_gemv_and_timestep_embedding_kernels(nat, dbl):
    G.test_gemv_and_timestep(nat, dbl)


def test_unet_input_cfg_ddim_and_table_kernels(nat, dbl):
    G.test_unet_input_cfg_ddim_tables(nat, dbl)


@pytest.mark.parametrize("N,D,q_first,q_count", [(3, 3, 1, 2)])   # D = 3 keys per pixel, a strict subset of query views (the sharded form)
def test_gridattn_kernels(nat, dbl, N, D, q_first, q_count):
    """rows a8 / a9 / a10 (per-op form) / a11's pyramid: prep, token producer, view attention, view pooling, frustum pooling, pixel
    cross-attention — the geometry (unproject, reproject, border-clamped bilinear taps, Plucker / depth harmonics) runs from the .cu source"""
    G.test_gridattn_kernels(nat, dbl, N, D, q_first, q_count)


@pytest.mark.parametrize("P,V", [(130, 8), (50, 3)])
def test_view_attention_kernels(nat, dbl, P, V):
    """staged kernel (heads * V divides 256, ragged last CTA) and the per-thread variant (V = 3)"""
    G.test_view_attention_layouts(nat, dbl, P, V)


@pytest.mark.parametrize("n,hw,C,silu", [(1, 1024, 320, True), (2, 64, 1280, True), (1, 100, 64, True), (1, 4096, 64, False), (2, 16, 2560, True)])
def test_groupnorm_cluster_kernel(nat, dbl, n, hw, C, silu):
    """gn_cluster_kernel: slices of a few groups, pixels interleaved over a cluster of 1-8 blocks, fp64 group sums exchanged through the
    peers' shared memory (32^2 x 320: a cluster split; 64^2: pixels streamed twice; ragged hw = 100)"""
    G.test_groupnorm(nat, dbl, n, hw, C, silu)


def test_groupnorm_two_sources_and_split_precision(nat, dbl):
    G.test_groupnorm_two_sources_and_concat16(nat, dbl, 1, 256, 640, 320)
    G.test_groupnorm_split_precision_output(nat, dbl, 1, 1024, 320)


def test_layernorm_family_and_softmax(nat, dbl):
    G.test_layernorms(nat, dbl)
    G.test_layernorm_f32_with_row_pitch(nat, dbl)
    t = {"s": G.rnd(70, 512) * 3, "p": G.torch.zeros(70, 512, dtype=G.torch.float16)}
    G.run_both(nat, dbl, "softmax_rows", t, ["p"], "s", "p", 70, 512, 512 ** -0.5)


# ------------------------------------------------------------------------------------------------ the whole step program, SIMT kernels from source
SIMT_OPS = ("groupnorm", "groupnorm_hilo", "groupnorm2", "layernorm", "layernorm_f32", "ln_modulate", "softmax_rows", "cast", "concat", "concat16",
            "upsample2x", "im2col_s2", "gemv", "gemv_grouped", "timestep_embedding", "unet_input", "cfg_ddim", "nchw_to_rows", "rows_to_nchw",
            "nchw_to_nhwc16", "gather_rows", "increment", "gridattn_prep", "gridattn_tokens", "view_attention", "view_pool", "frustum_pool",
            "pixel_cross_attn")


@pytest.mark.skipif(os.environ.get("MVD_SLOW_SHIM") != "1",
                    reason="two minutes (emulated GEMMs + 1.5 M fiber switches): run with MVD_SLOW_SHIM=1 (last result: profiles/r02_shim_apply_model.txt)")
def test_apply_model_with_every_simt_kernel_from_source_vs_oracle(nat, monkeypatch):
    """ViewFusion.apply_model (N = 2, D = 3, cfg 2.5, the 64-channel topology-complete model) through the product's step compiler with EVERY
    non-tensor-core kernel executed from its .cu source on the shim — in the real program: the engine's buffer arena, pitches, column
    windows, two-source GroupNorms, CLIP row biases — and only the tcgen05 kernels (GEMM / implicit-GEMM conv, flash attention, fused DiT)
    emulated; against the fp32 oracle at the path's own gate."""
    import torch
    import mvdfusion_b200.runtime as rt
    from common import build_model, rel_l2, state_dict_cpu, unet_cfg_of
    from test_host_engine import cams_of
    from mvdfusion_b200 import synthetic
    from ops_double import TorchOpsDouble
    from oracle import mvd_oracle as O

    class HybridOps(TorchOpsDouble):
        used = set()

        def __getattribute__(self, name):
            if name in SIMT_OPS:
                HybridOps.used.add(name)
                return getattr(nat, name)
            return object.__getattribute__(self, name)

    hyb = HybridOps()
    monkeypatch.setattr(rt, "get_ops", lambda dev: hyb)
    D = 3
    m = build_model(64, 8, D=D, S=32)
    sd = state_dict_cpu(m)
    sc = synthetic.scene_inputs(2, 32)
    de, _ = synthetic.step_noises(2, D, 32, 1)
    t = torch.full((2,), 501, dtype=torch.long)
    eps = m.apply_model(sc["x_T"], cams_of(sc["cams"]), sc["input_latents"], cams_of(sc["in_cams"]), sc["clip_v_embed"], t, cfg_scale=2.5, depth_eps=de[0])
    ref = O.apply_model(sd, sc["x_T"], sc["cams"], sc["input_latents"], sc["in_cams"], sc["clip_v_embed"], t, de[0], unet_cfg=unet_cfg_of(m), D=D, cfg_scale=2.5)
    assert torch.isfinite(eps).all()
    r = rel_l2(eps, ref)
    print("apply_model, SIMT kernels from source: rel-L2 vs oracle", r, "ops from source:", sorted(HybridOps.used))
    assert r < 1e-3, r
    assert {"groupnorm", "gridattn_prep", "gridattn_tokens", "unet_input", "cfg_ddim", "gemv_grouped", "frustum_pool", "pixel_cross_attn"} <= HybridOps.used


@pytest.mark.parametrize("order", ["reverse", "shuffle:7"])
def test_results_do_not_depend_on_the_thread_schedule(nat, dbl, monkeypatch, order):
    """A CPU stand-in for racecheck: the shim resumes runnable threads in a different order between barriers (MVD_SHIM_ORDER); kernels
    whose shared-memory traffic is correctly fenced by __syncthreads / shuffles / cluster barriers give the same answers — here the
    cluster GroupNorm (partials in shared memory, DSMEM exchange), the grouped GEMV (staged input row), the staged view attention and
    the token producer, against the same references at the same tolerances."""
    monkeypatch.setenv("MVD_SHIM_ORDER", order)
    G.test_groupnorm(nat, dbl, 1, 1024, 320, True)
    G.test_groupnorm_two_sources_and_concat16(nat, dbl, 1, 256, 640, 320)
    G.test_gemv_grouped(nat, dbl)
    G.test_view_attention_layouts(nat, dbl, 130, 8)
    G.test_layernorms(nat, dbl)
