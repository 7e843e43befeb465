"""CPU: the SOURCE of the plain-SIMT kernels of the hot path (csrc/elementwise.cu: cast / concat / upsample / im2col / GEMV / grouped GEMV /
timestep embedding / UNet input assembly / CFG combine + DDIM update / layout converters / table gather) compiled as C++ and executed on
host threads (tests/native/cpu_emul/cuda_on_cpu.h), through the product's own ops.NativeOps bindings, against the same emulations of the
documented semantics the B200 tests use — the very test bodies of tests/test_gpu_ops.py, with `nat` bound to the shim build.

What this proves without a GPU: indexing, tails, fp16 rounding (IEEE binary16 through _Float16), the grouped-GEMV job table, the DDIM /
CFG arithmetic.  What it cannot prove: anything about tcgen05 / TMA / clusters (gemm.cu, attention.cu, dit.cu, norm.cu's cluster
GroupNorm are hardware-only) or timing; the -m gpu suite remains the proof on the B200."""
import pytest

import test_gpu_ops as G
from common import build_cpu_shim, shim_ops


@pytest.fixture(scope="module")
def shim_lib(tmp_path_factory):
    return build_cpu_shim(["elementwise.cu"], tmp_path_factory.mktemp("cpu_shim"), "libmvd_elementwise_cpuemul.so")


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
