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


@pytest.mark.parametrize("shape,ndist", [("32,32,32", 19), ("24,16,16", 27), ("64,64", 9), ("64,32,32", 19)])
def test_partitioned_gpu_peer_to_peer_halo(shape, ndist):
    """The peer-to-peer halo (lbm_b200_p2p_export / _import: CUDA IPC mailboxes, device-to-device copies on the copy engines, flag words,
    ONE launch per step with the outer tiles first) instead of ncclSend / ncclRecv: owned cells bit-identical to the single-domain oracle,
    and the all-reduced residual (still NCCL) agrees."""
    _need_two_gpus()
    launch(2, "gpu", shape, ndist, steps=14, extra=("--p2p", "--check-residual"))


@pytest.mark.parametrize("case,collision,level", [("step3d", "trt", 6)])
def test_partitioned_gpu_peer_to_peer_halo_with_pressure_boundary(case, collision, level):
    _need_two_gpus()
    launch(2, "gpu", "0,0,0", 0, steps=10, extra=("--case", case, "--level", str(level), "--collision", collision, "--p2p"))


def test_cpp_host_two_ranks_match_one_rank(tmp_path):
    """`lbm` as two processes (one GPU each; rank / world size from the environment, NCCL id through a file): the moments each rank
    writes for its own cells must equal the single-process run's, bit for bit (STRICT fp64).  The set-up half is pinned on the CPU
    (tests/test_host_partitioned.py); this is the NCCL bootstrap and the run loop."""
    import base64
    import json
    import os
    import re
    import subprocess

    import numpy as np

    from lbm_b200 import cases
    _need_two_gpus()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "lbm_b200", "lbm")
    cfg = cases.CONFIGS["step3d"](5)
    cfg["solver"]["maxSteps"] = 40
    cfg["solver"]["solution_interval"] = 10 ** 9
    cfg["solver"]["info_interval"] = 10
    path = tmp_path / "case.json"
    path.write_text(json.dumps(cfg))

    def rho_of(file):
        text = open(file).read()
        m = re.search(r'Name="rho"[^>]*>\s*\n([A-Za-z0-9+/=]*)\n', text)
        raw = base64.b64decode(m.group(1) + "=" * (-len(m.group(1)) % 4))
        return np.frombuffer(raw[8:8 + (len(raw) - 8) // 8 * 8], dtype=np.float64)

    one = tmp_path / "one"
    one.mkdir()
    r = subprocess.run([exe, str(path)], cwd=one, capture_output=True, text=True, timeout=600,
                       env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
    assert r.returncode == 0, r.stderr[-2000:]
    ref = rho_of(one / "out" / "solution_39.vtp")
    two = tmp_path / "two"
    two.mkdir()
    procs = []
    for rank in range(2):
        env = dict(os.environ, LBM_B200_RANK=str(rank), LBM_B200_WORLD="2", LBM_B200_LOCAL_RANK=str(rank), LBM_B200_ID_FILE=str(two / "nccl_id"))
        procs.append(subprocess.Popen([exe, str(path)], cwd=two, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    parts = [rho_of(two / "out" / f"solution_39_rank{rank}.vtp") for rank in range(2)]
    n = len(ref)
    assert len(parts[0]) == n // 2 and len(parts[0]) + len(parts[1]) == n
    # the solution file stores values rounded to 15 decimals (the reference's writer): compare what both runs wrote
    assert np.array_equal(np.concatenate(parts), ref)
