"""NCCL halves of the partition tests added late in round 1: WRITTEN AND PINNED ON CPU (gloo oracle runs in tests/test_distributed.py,
plan interpreter in tests/test_plan_cpu.py and tests/test_baseline_configs.py) but NOT YET EXECUTED OVER NCCL -- the round's GPU budget
was spent; they need >= 2 GPUs.  Kept in a file that sorts last so that a surprise here cannot hide the verified GPU tests behind -x.
  * velocity halo of the pressure boundary condition (lbm_b200_set_vars_halo, k_velocity_pack)
  * residual of the whole domain (ncclAllReduce inside lbm_b200_residual)
  * BASELINE.json configs[3] / [4] (3D sphere D3Q27 MRT, 3D step D3Q19 TRT) cut into two SFC ranges"""
import pytest

from test_distributed import launch

pytestmark = pytest.mark.gpu


def _need_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


@pytest.mark.parametrize("shape,ndist", [("18,16,16", 19), ("17,16,16", 27), ("34,64", 9)])
def test_partitioned_gpu_pressure_boundary_nccl(shape, ndist):
    _need_two_gpus()
    launch(2, "gpu", shape, ndist, bc="pressure", extra=("--check-residual",))


@pytest.mark.parametrize("shape,ndist", [("32,32,32", 19)])
def test_partitioned_gpu_residual_allreduce_nccl(shape, ndist):
    _need_two_gpus()
    launch(2, "gpu", shape, ndist, extra=("--check-residual",))


@pytest.mark.parametrize("case,collision,level", [("sphere3d", "mrt", 6), ("step3d", "trt", 6)])
def test_partitioned_gpu_baseline_configs_nccl(case, collision, level):
    _need_two_gpus()
    launch(2, "gpu", "0,0,0", 0, steps=10, extra=("--case", case, "--level", str(level), "--collision", collision))
