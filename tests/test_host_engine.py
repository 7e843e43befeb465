"""CPU: the product's host-side logic (reference-surface modules -> engine emitters -> program of bound calls, weight
packing, buffer arena, step tables) executed against the CPU emulation of the C ABI and compared with the oracle.
What this does NOT test is the kernels: that is tests/test_gpu_*.py on the B200."""
import pytest
import torch

from common import build_model, rel_l2, state_dict_cpu, synthetic, unet_cfg_of
from mvdfusion_b200.mvdfusion.cameras import PerspectiveCameras
from oracle import mvd_oracle as O

TOL = 2e-3  # the emulation rounds fp16 tensors exactly like the kernels' outputs


def cams_of(c):
    return PerspectiveCameras(c["R"], c["T"], c["f"], c["p"])


@pytest.fixture(scope="module")
def model():
    m = build_model(64, 8, D=1, S=32)
    return m, state_dict_cpu(m)


def test_state_dict_names_follow_reference(model):
    m, sd = model
    for k in ("view_attn.z_embedder.0.weight", "view_attn.t_embedder.mlp.0.weight", "view_attn.pre_layer_b.0.weight",
              "view_attn.aggregation_transformer.layer_list.2.attn.qkv.bias",
              "view_attn.aggregation_transformer.layer_list.0.adaLN_modulation.1.weight",
              "view_attn.aggregation_transformer.weight_layer.weight", "view_attn.final_layer_b.weight",
              "unet_model.unet_model.time_embed.0.weight", "unet_model.unet_model.input_blocks.0.0.weight",
              "unet_model.unet_model.input_blocks.1.1.transformer_blocks.0.attn1.to_q.weight",
              "unet_model.unet_model.input_blocks.3.0.op.weight", "unet_model.unet_model.middle_block.2.aligned_attn_proj_in.weight",
              "unet_model.unet_model.middle_block.2.aligned_attn_transformer_blocks.0.attn2.to_k.weight",
              "unet_model.unet_model.output_blocks.5.3.conv.weight", "unet_model.unet_model.output_blocks.11.2.aligned_attn_norm.weight",
              "unet_model.unet_model.output_blocks.2.1.conv.weight", "unet_model.unet_model.out.2.bias",
              "scheduler.alphas_cumprod", "cc_projection.4.weight", "time_embed.2.bias"):
        assert k in sd, k
    assert sd["view_attn.pre_layer_b.0.weight"].shape == (256, 723)
    assert sd["cc_projection.0.weight"].shape == (768, 796)


@pytest.mark.parametrize("D,cfg", [(1, 2.5), (3, 2.5), (1, 1.0)])
def test_apply_model(ops_double, D, cfg):
    m = build_model(64, 8, D=D, S=32)
    sd = state_dict_cpu(m)
    sc = synthetic.scene_inputs(2, 32)
    de, _ = synthetic.step_noises(2, D, 32, 1)
    t = torch.full((2,), 501, dtype=torch.long)
    eps = m.apply_model(sc["x_T"], cams_of(sc["cams"]), sc["input_latents"], cams_of(sc["in_cams"]), sc["clip_v_embed"], t,
                        cfg_scale=cfg, depth_eps=de[0])
    ref = O.apply_model(sd, sc["x_T"], sc["cams"], sc["input_latents"], sc["in_cams"], sc["clip_v_embed"], t, de[0],
                        unet_cfg=unet_cfg_of(m), D=D, cfg_scale=cfg)
    assert torch.isfinite(eps).all()
    assert rel_l2(eps, ref) < TOL


def test_apply_model_condition_drop_and_prev_depth(ops_double, model):
    m, sd = model
    sc = synthetic.scene_inputs(2, 32)
    de, _ = synthetic.step_noises(2, 1, 32, 1)
    t = torch.full((2,), 501, dtype=torch.long)
    rnd = torch.tensor([0.03, 0.12])
    prev = torch.rand(2, 1, 32, 32) * 2 - 1
    m.drop_conditions = True
    try:
        eps = m.apply_model(sc["x_T"], cams_of(sc["cams"]), sc["input_latents"], cams_of(sc["in_cams"]), sc["clip_v_embed"], t,
                            cfg_scale=1.0, depth_eps=de[0], drop_random=rnd, prev_depth=prev)
    finally:
        m.drop_conditions = False
    ref = O.apply_model(sd, sc["x_T"], sc["cams"], sc["input_latents"], sc["in_cams"], sc["clip_v_embed"], t, de[0],
                        unet_cfg=unet_cfg_of(m), D=1, cfg_scale=1.0, drop_random=rnd, prev_depth=prev)
    assert rel_l2(eps, ref) < TOL


def test_ddim_loop_tables_and_counter(ops_double, model):
    m, sd = model
    steps = 4  # (1000 must be divisible: the reference has the same constraint, util.py:48-49)
    m.ddim._make_schedule(steps, "uniform", 1.0)
    sc = synthetic.scene_inputs(2, 32)
    de, dn = synthetic.step_noises(2, 1, 32, steps)
    x, inter = m.ddim.sample(cams_of(sc["cams"]), sc["input_latents"], cams_of(sc["in_cams"]), sc["clip_v_embed"],
                             unconditional_scale=2.5, depth=True, return_intermediates=True, verbose=False, x_T=sc["x_T"],
                             depth_eps=de, ddim_noise=dn)
    ref, rinter = O.ddim_sample(sd, sc["x_T"], sc["cams"], sc["input_latents"], sc["in_cams"], sc["clip_v_embed"], de, dn,
                                unet_cfg=unet_cfg_of(m), D=1, num_steps=steps, eta=1.0, cfg_scale=2.5, return_intermediates=True)
    assert [a["t"] for a in inter] == [b["t"] for b in rinter]
    for a, b in zip(inter, rinter):
        assert rel_l2(a["xt"], b["xt"]) < TOL and rel_l2(a["x0"], b["x0"]) < TOL
    assert rel_l2(x, ref) < TOL


def test_module_level_forwards(ops_double, model):
    m, sd = model
    um = m.unet_model.unet_model
    pre = "unet_model.unet_model."
    g = torch.Generator().manual_seed(3)
    # ResBlock with channel change (1x1 skip) and per-image embeddings
    rb = um.input_blocks[4][0]
    x = torch.randn(2, 64, 16, 16, generator=g)
    emb = torch.randn(2, 256, generator=g)
    assert rel_l2(rb(x, emb), O.resblock(sd, pre + "input_blocks.4.0", x, emb)) < TOL
    # SpatialTransformer
    st = um.input_blocks[1][1]
    x = torch.randn(2, 64, 32, 32, generator=g)
    ctx = torch.randn(2, 1, 768, generator=g)
    assert rel_l2(st(x, ctx), O.spatial_transformer(sd, pre + "input_blocks.1.1", x, ctx, 8)) < TOL
    # ViewAlignedFeatureTransformer at the 8x8 level with D = 2 keys per pixel
    va = um.output_blocks[3][2]
    x = torch.randn(2, 256, 8, 8, generator=g)
    pyr = [torch.randn(2, 32 >> l, 32 >> l, 2, 768, generator=g) for l in range(4)]
    assert rel_l2(va(x, pyr), O.view_aligned_transformer(sd, pre + "output_blocks.3.2", x, pyr, 8, 32)) < TOL
    # self-attention
    ca = st.transformer_blocks[0].attn1
    x = torch.randn(2, 256, 64, generator=g)
    assert rel_l2(ca(x), O.cross_attention(sd, pre + "input_blocks.1.1.transformer_blocks.0.attn1", x, None, 8)) < TOL
    # whole UNet through its own forward
    xin = torch.randn(2, 10, 32, 32, generator=g)
    vol = torch.randn(2, 32, 32, 1, 768, generator=g)
    y = um(xin, torch.tensor([301]), ctx, volume_feats=O.volume_pyramid(vol))
    ref = O.unet_forward(sd, xin, torch.tensor([301]), ctx, O.volume_pyramid(vol), model_channels=64, num_heads=8, image_size=32, prefix=pre)
    assert rel_l2(y, ref) < TOL


def test_step_program_has_no_cast_or_concat_passes(ops_double, model, monkeypatch):
    """The fp16 operands of the 1x1 skip convolutions are written by their producers' epilogues (mvd_gemm_args.out16 windows of
    one [rows, c1+c2] buffer per output block): the step program contains no cast / concat launch, the MVD_NO_FUSE_CAT path does,
    and both give the same result."""
    import mvdfusion_b200.engine as E
    m, _ = model
    sc = synthetic.scene_inputs(2, 32)
    de, _ = synthetic.step_noises(2, 1, 32, 1)
    t = torch.full((2,), 501, dtype=torch.long)
    args = (sc["x_T"], cams_of(sc["cams"]), sc["input_latents"], cams_of(sc["in_cams"]), sc["clip_v_embed"], t)

    def names(plan):
        return [getattr(c, "__name__", None) or getattr(c, "name", "") for c in plan.core_prog.calls]

    # MVD_HILO=1: stem / head split precision only — the [hi | lo] operand of the split-precision skip convolutions (level 2)
    # exists only in the producer-written form, so the bit-for-bit comparison is made one level down
    monkeypatch.setenv("MVD_HILO", "1")
    m = build_model(64, 8, D=1, S=32)
    eps_fused = m.apply_model(*args, cfg_scale=2.5, depth_eps=de[0])
    plan = m.step_plan(2, 32, 1, use_cfg=True)
    n_fused = len(plan.core_prog)
    monkeypatch.setenv("MVD_NO_FUSE_CAT", "1")
    m2 = build_model(64, 8, D=1, S=32)
    eps_plain = m2.apply_model(*args, cfg_scale=2.5, depth_eps=de[0])
    n_plain = len(m2.step_plan(2, 32, 1, use_cfg=True).core_prog)
    # 12 skip concatenations + 2 casts in front of the channel-changing input ResBlocks + 3 im2col passes of the Downsample convolutions
    # (strided implicit GEMM on the producer-written fp16 operand)
    assert n_plain - n_fused == 17
    assert torch.equal(eps_fused, eps_plain)


@pytest.mark.parametrize("D", [1, 3])
def test_step_program_has_no_layernorm_passes(ops_double, D, monkeypatch):
    """norm1 / norm3 of every transformer block (SpatialTransformer and ViewAlignedFeatureTransformer) ride in their neighbours'
    epilogues (mvd_gemm_args.ln_stats_out / ln_stats, ABI 13): the step program holds no LayerNorm launch for them; the MVD_NO_LN_FOLD
    program (one ln_kernel pass each) is equally close to the fp32 oracle."""
    sc = synthetic.scene_inputs(2, 32)
    de, _ = synthetic.step_noises(2, D, 32, 1)
    t = torch.full((2,), 501, dtype=torch.long)
    args = (sc["x_T"], cams_of(sc["cams"]), sc["input_latents"], cams_of(sc["in_cams"]), sc["clip_v_embed"], t)

    def count(plan, name):
        return sum(1 for c in plan.core_prog.calls if getattr(c, "name", "") == name)

    m = build_model(64, 8, D=D, S=32)
    eps_fold = m.apply_model(*args, cfg_scale=2.5, depth_eps=de[0])
    plan = m.step_plan(2, 32, D, use_cfg=True)
    monkeypatch.setenv("MVD_NO_LN_FOLD", "1")
    m2 = build_model(64, 8, D=D, S=32)
    eps_plain = m2.apply_model(*args, cfg_scale=2.5, depth_eps=de[0])
    plan2 = m2.step_plan(2, 32, D, use_cfg=True)
    n_fold, n_plain = count(plan, "layernorm"), count(plan2, "layernorm")
    blocks = n_plain // (2 if D == 1 else 2.5)  # D > 1: norm2 of the view-aligned blocks keeps its pass (its consumer is to_q)
    assert n_plain > 0 and n_plain - n_fold >= 2 * int(blocks) - 1 and (n_fold == 0 if D == 1 else n_fold < n_plain)
    assert len(plan2.core_prog) - len(plan.core_prog) == n_plain - n_fold
    # two different fp16 rounding patterns of the same arithmetic: each is ~7e-4 from the fp32 oracle, so they sit ~7e-4 apart;
    # what counts is that the folded program is as close to the oracle as the LayerNorm-pass one
    sd = state_dict_cpu(m)
    ref = O.apply_model(sd, sc["x_T"], sc["cams"], sc["input_latents"], sc["in_cams"], sc["clip_v_embed"], t, de[0],
                        unet_cfg=unet_cfg_of(m), D=D, cfg_scale=2.5)
    e_fold, e_plain = rel_l2(eps_fold, ref), rel_l2(eps_plain, ref)
    assert e_fold < TOL and e_fold < 1.1 * e_plain + 1e-5
    assert rel_l2(eps_fold, eps_plain) < 2 * TOL


def test_step_program_has_no_upsample_passes(ops_double, monkeypatch):
    """Upsample (nearest x2 + conv3x3, openaimodel.py:107-119) runs as four 2 x 2 phase convolutions of the source image in one launch
    (mvd_gemm_args.conv_up2, ABI 14): no upsample2x launch and no upsampled tensor in the step program; the MVD_NO_FOLD_UP program is
    equally close to the fp32 oracle."""
    sc = synthetic.scene_inputs(2, 32)
    de, _ = synthetic.step_noises(2, 1, 32, 1)
    t = torch.full((2,), 501, dtype=torch.long)
    args = (sc["x_T"], cams_of(sc["cams"]), sc["input_latents"], cams_of(sc["in_cams"]), sc["clip_v_embed"], t)

    def count(plan, name):
        return sum(1 for c in plan.core_prog.calls if getattr(c, "name", "") == name)

    m = build_model(64, 8, D=1, S=32)
    eps_fold = m.apply_model(*args, cfg_scale=2.5, depth_eps=de[0])
    plan = m.step_plan(2, 32, 1, use_cfg=True)
    monkeypatch.setenv("MVD_NO_FOLD_UP", "1")
    m2 = build_model(64, 8, D=1, S=32)
    eps_plain = m2.apply_model(*args, cfg_scale=2.5, depth_eps=de[0])
    plan2 = m2.step_plan(2, 32, 1, use_cfg=True)
    assert count(plan, "upsample2x") == 0 and count(plan2, "upsample2x") == 3 and count(plan, "_gemm_up2") == 3
    assert len(plan2.core_prog) - len(plan.core_prog) == 3
    sd = state_dict_cpu(m)
    ref = O.apply_model(sd, sc["x_T"], sc["cams"], sc["input_latents"], sc["in_cams"], sc["clip_v_embed"], t, de[0],
                        unet_cfg=unet_cfg_of(m), D=1, cfg_scale=2.5)
    e_fold, e_plain = rel_l2(eps_fold, ref), rel_l2(eps_plain, ref)
    assert e_fold < TOL and e_fold < 1.1 * e_plain + 1e-5


def test_weight_cache_follows_parameter_updates(ops_double, model):
    m, sd = model
    st = m.unet_model.unet_model.input_blocks[1][1]
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 64, 32, 32, generator=g)
    ctx = torch.randn(1, 1, 768, generator=g)
    y0 = st(x, ctx)
    with torch.no_grad():
        st.proj_out.weight.mul_(2.0)
    y1 = st(x, ctx)
    with torch.no_grad():
        st.proj_out.weight.mul_(0.5)
    assert rel_l2(y1 - x, 2 * (y0 - x) - st.proj_out.bias.view(1, -1, 1, 1)) < 1e-2  # packed weights were rebuilt


def test_prepare_batch_matches_reference_layout(model):
    m, _ = model
    R, T, f, p = synthetic.gso_rig(8)
    # un-relative rig: rotate the world so that view 0 is NOT the identity, prepare_batch must undo it
    Q = torch.linalg.qr(torch.randn(3, 3, generator=torch.Generator().manual_seed(0)))[0]
    batch = {"latents": torch.randn(9, 4, 32, 32), "clip_embed": torch.randn(9, 1, 768), "R": torch.einsum("ij,bjk->bik", Q, R),
             "T": T, "f": f, "c": p}
    cfg = {"input_batch_size": 1, "train_batch_size": 8, "random_views": False}
    bl, bc, il, ic, cve = m.prepare_batch(batch, cfg)
    assert bl.shape == (8, 5, 32, 32) and il.shape == (1, 5, 32, 32) and cve.shape == (8, 1, 796)
    assert float(il[:, 4].abs().max()) == 0.0                         # input depth is zeroed (viewfusion…py:215)
    assert torch.allclose(ic.R[0], torch.eye(3), atol=1e-5)           # relative to the input view
    assert torch.allclose(bc.R, R[1:], atol=1e-5)
    assert torch.allclose(cve[0, 0, 768:777], ic.R.reshape(-1), atol=1e-6)
    assert torch.allclose(cve[3, 0, 782:791], bc.R[3].reshape(-1), atol=1e-6)
    assert torch.allclose(cve[3, 0, 794:796], bc.focal_length[3], atol=1e-6)


def test_sample_and_forward_through_the_facade_baseline_config0(ops_double):
    """BASELINE.json configs[0] plumbing (demo.py:85 -> ViewFusion.sample): one input view, N = 4 target views, 10 DDIM steps,
    cfg 2.5, random-init (64-channel) UNet; the facade draws x_T and the per-step noises itself, in the reference's order
    (sampler.py:107,128 / view_attn_efficient2.py:431), so seeding torch reproduces them for the oracle.  Also the training
    facade: ViewFusion.forward(batch, cfg) = mean squared error of the predicted noise (viewfusion_zero_depth_rgb.py:362-392)."""
    N, S, steps = 4, 32, 10
    m = build_model(64, 8, D=1, S=S, ddim_steps=steps)
    sd = state_dict_cpu(m)
    R, T, f, p = synthetic.gso_rig(N)
    g = torch.Generator().manual_seed(11)
    batch = {"latents": torch.randn(N + 1, 4, S, S, generator=g) * 0.8, "clip_embed": torch.randn(N + 1, 1, 768, generator=g),
             "R": R, "T": T, "f": f, "c": p}
    cfg = {"input_batch_size": 1, "train_batch_size": N, "random_views": False}
    torch.manual_seed(123)
    x, bl, il, bc, inter = m.sample(batch, cfg, cfg_scale=2.5, return_input=True, depth=True, verbose=False)
    assert x.shape == (N, 5, S, S) and len(inter) == steps and bl.shape == (N, 5, S, S)
    # the same draws for the oracle
    torch.manual_seed(123)
    x_T = torch.randn([N, 5, S, S])
    de, dn = [], []
    for _ in range(steps):
        de.append(torch.randn(N, 1, S, S))
        dn.append(torch.randn(N, 5, S, S))
    _, bc2, il2, ic2, cve = m.prepare_batch(batch, cfg)
    cams = {"R": bc2.R, "T": bc2.T, "f": bc2.focal_length, "p": bc2.principal_point}
    icams = {"R": ic2.R, "T": ic2.T, "f": ic2.focal_length, "p": ic2.principal_point}
    ref, _ = O.ddim_sample(sd, x_T, cams, il2, icams, cve, torch.stack(de), torch.stack(dn), unet_cfg=unet_cfg_of(m), D=1,
                           num_steps=steps, eta=1.0, cfg_scale=2.5, return_intermediates=True)
    assert rel_l2(x, ref) < 3 * TOL  # ten chained steps

    torch.manual_seed(7)
    loss = m(batch, cfg)
    assert loss.dim() == 0 and torch.isfinite(loss) and 0.1 < float(loss) < 10.0


def test_plan_cache_is_bounded_lru():
    """plans own arenas / graphs; the cache keeps the most recently used few"""
    from mvdfusion_b200.runtime import PlanCache
    c = PlanCache(3)
    for k in "abc":
        c[k] = k.upper()
    assert c["a"] == "A"          # touch: 'b' is now the oldest
    c["d"] = "D"
    assert list(c) == ["c", "a", "d"] and "b" not in c
    c["c"] = "C2"
    c["e"] = "E"
    assert list(c) == ["d", "c", "e"] and c["c"] == "C2"
