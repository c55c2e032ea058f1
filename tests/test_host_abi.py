"""CPU-side checks: the C-ABI library loads and exports exactly what include/so100_b200.h declares,
fails loudly without a GPU (no fallback), and the host-side logic (spaces, seeding, sharding, the
2-rank gloo statistics all-reduce) behaves."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "so100_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(so100_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gym_so100_c_b200 import build, ext
    build.build()                        # nvcc cross-compiles sm_100a without a GPU
    lib = C.CDLL(build.LIB)
    declared = _declared_symbols()
    assert set(declared) == set(ext.SYMBOLS), (declared, ext.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_library_contains_sm100a_code_only():
    from gym_so100_c_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_model_header_matches_dtype(model_blob):
    from gym_so100_c_b200 import model
    header = open(os.path.join(ROOT, "include", "so100_model.h")).read()
    assert header == model.c_header()
    assert f"#define SO100_MODEL_BYTES {len(model_blob)}" in header
    m = model.unpack(model_blob)
    assert int(m["magic"]) == model.MAGIC
    with pytest.raises(ValueError):
        model.unpack(model_blob[:-8])


def test_create_fails_loudly_without_gpu(model_blob):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gym_so100_c_b200 import ext
    lib = ext.load()
    h = C.c_void_p()
    rc = lib.so100_create(model_blob, len(model_blob), 4, 0, 0, C.c_uint64(0), C.c_int64(0), C.byref(h))
    assert rc == -2 and b"no CPU path" in lib.so100_last_error()
    rc = lib.so100_create(model_blob[:100], 100, 4, 0, 0, C.c_uint64(0), C.c_int64(0), C.byref(h))
    assert rc == -1
    from gym_so100_c_b200.vec_env import SO100VecEnv
    with pytest.raises(ext.So100Error):
        SO100VecEnv(4)
    with pytest.raises(ext.So100Error):
        SO100VecEnv(4, task="so100_touch_cube")          # a registered task: fails only because there is no GPU
    with pytest.raises(NotImplementedError):
        SO100VecEnv(4, task="so100_transfer_cube")       # env.py:117-118
    rc = lib.so100_create(model_blob, len(model_blob), 4, 0, 7, C.c_uint64(0), C.c_int64(0), C.byref(h))
    assert rc == -1 and b"unknown task" in lib.so100_last_error()
    with pytest.raises(NotImplementedError):
        SO100VecEnv(4, obs_type="so100_pixels")               # not one of env.py:50-73's two observation types


def test_spaces_shim():
    from gym_so100_c_b200.spaces import Box, Dict, batch_box
    a = Box(low=-1, high=1, shape=(6,), dtype=np.float32)
    assert a.shape == (6,) and a.contains(a.sample()) and not a.contains(np.full(6, 2, np.float32))
    b = batch_box(a, 5)
    assert b.shape == (5, 6)
    d = Dict({"x": a})
    assert d.contains({"x": a.sample()})


def test_shard_ranges_partition():
    from gym_so100_c_b200.parallel import shard_range
    for n in (1, 7, 16384, 1048576, 1000003):
        for world in (1, 2, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from gym_so100_c_b200 import parallel
rank, world, local = parallel.init_from_env(backend="gloo")
lo, hi = parallel.shard_range(1001, rank, world)
stats = {k: (rank + 1) * (i + 1) for i, k in enumerate(parallel.STAT_KEYS)}
stats["episodes"] = hi - lo
tot = parallel.all_reduce_stats(stats)
mx = parallel.max_over_ranks(10.0 + rank)
parallel.barrier()
assert tot["episodes"] == 1001, tot
assert tot["contact_overflow"] == sum(range(1, world + 1)), tot
assert mx == 10.0 + world - 1, mx
ep = parallel.all_reduce_episode_stats({"episodes": 10 * (rank + 1), "successes": rank, "return_sum": -5.0 * (rank + 1), "length_sum": 100 * (rank + 1)})
assert ep["episodes"] == 30 and ep["successes"] == 1 and ep["return_sum"] == -15.0 and ep["length_sum"] == 300, ep
assert abs(ep["ep_rew_mean"] + 0.5) < 1e-12 and abs(ep["ep_len_mean"] - 10.0) < 1e-12 and abs(ep["success_rate"] - 1 / 30) < 1e-12, ep
assert parallel.gather_floats(1.5 + rank) == [1.5 + r for r in range(world)]
if rank == 0:
    print("GLOO_OK", world, tot["episodes"])
dist.destroy_process_group()
"""


def test_two_rank_gloo_statistics(tmp_path):
    """world_size-2 gloo run of the N>1 host path: shard ranges + episode-statistics all-reduce + max-over-ranks."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613", str(script), ROOT],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_OK 2 1001" in out.stdout


def test_demonstration_loading_and_start_poses(tmp_path):
    """The reference's recording format (scripts/record_teleop.py:177-184, 277-282): a pickled list of episode dicts.
    State observations give the cube's start pose (cube_site - 0.01), pixel observations fall back to the seeded pose."""
    import pickle
    import numpy as np
    from gym_so100_c_b200 import replay
    from gym_so100_c_b200.vec_env import sample_so100_box_pose
    obs0 = np.zeros(15, np.float32); obs0[:3] = [-0.19, 0.46, 0.06]
    eps = [dict(observations=[obs0, obs0], actions=[np.zeros(6), np.ones(6) * 0.5], rewards=[0.0, 1.0], infos=[{}, {}]),
           dict(observations=[{"pixels": None, "agent_pos": np.zeros(6)}], actions=[np.zeros(6)], rewards=[0.0], infos=[{}])]
    path = tmp_path / "expert_demonstrations.pkl"
    with open(path, "wb") as f:
        pickle.dump(eps, f)
    loaded = replay.load_demonstrations(str(path))
    assert len(loaded) == 2 and len(loaded[0]["actions"]) == 2
    poses = replay.start_poses(loaded, seed=5)
    assert np.allclose(poses[0], [-0.2, 0.45, 0.05, 1, 0, 0, 0], atol=1e-6)
    assert np.array_equal(poses[1], sample_so100_box_pose(6).astype(np.float32))
    with open(path, "wb") as f:
        pickle.dump([dict(observations=[])], f)
    with pytest.raises(ValueError):
        replay.load_demonstrations(str(path))


def test_lazy_infos_behave_like_a_list_of_dicts():
    """The SB3 adapter's infos (vec_env._LazyInfos): list protocol, keys of the reference's stack, terminal entries only for
    envs that finished."""
    import numpy as np
    from gym_so100_c_b200.vec_env import _LazyInfos
    n = 5
    done = np.array([0, 1, 0, 1, 0], bool)
    final = np.arange(n * 15, dtype=np.float32).reshape(n, 15)
    prev = np.arange(n * 3, dtype=np.float32).reshape(n, 3)
    infos = _LazyInfos(succ=done.copy(), timeout=~done, done=done, final=final, goal_env=True, prev_desired=prev,
                       ep_return=np.arange(n, dtype=np.float32), ep_length=np.arange(n, dtype=np.int32) * 10)
    assert len(infos) == n and len(list(infos)) == n and len(infos[1:3]) == 2
    assert set(infos[0]) == {"is_success", "TimeLimit.truncated"} and infos[0]["TimeLimit.truncated"] is True
    t = infos[3]
    assert t["is_success"] is True and t["episode"] == {"r": 3.0, "l": 30}
    assert np.array_equal(t["terminal_observation"]["observation"], final[3]) and np.array_equal(t["terminal_observation"]["desired_goal"], prev[3])
    assert np.array_equal(infos[-1 - 1]["terminal_observation"]["achieved_goal"], final[3, :3])
    with pytest.raises(IndexError):
        infos[n]
    flat = _LazyInfos(succ=done, timeout=done, done=done, final=final, goal_env=False, prev_desired=None,
                      ep_return=np.zeros(n, np.float32), ep_length=np.zeros(n, np.int32))
    assert np.array_equal(flat[1]["terminal_observation"], final[1])


def test_render_tables_are_consistent(model_rec):
    """Renderer scene tables (render.scene_tables): every hull's facet planes contain all of its vertices and each is touched by
    at least three of them; pads are not drawn; the camera frames follow MuJoCo's targetbody rule."""
    import numpy as np
    from gym_so100_c_b200 import render
    m = model_rec
    planes, adr, num, rgb = render.scene_tables(m)
    assert (num[[g for g in range(int(m["ngeom"])) if (int(m["pad_mask"]) >> g) & 1]] == -1).all() and (num == -1).sum() == 8
    assert num[int(m["cg_cube"])] == 0 and num[int(m["cg_table"])] == 0 and tuple(rgb[int(m["cg_cube"])]) == (1.0, 0.0, 0.0)
    for g in range(int(m["ngeom"])):
        if num[g] <= 0:
            continue
        a = int(m["geom_vadr"][g])
        V = np.asarray(m["vert"][a:a + int(m["geom_vnum"][g])], float)
        P = planes[adr[g]:adr[g] + num[g]].astype(float)
        h = V @ P[:, :3].T - P[:, 3]                       # signed height of every vertex over every plane
        assert h.max() < 2e-6 and ((np.abs(h) < 2e-6).sum(axis=0) >= 3).all(), g
        assert np.abs(np.linalg.norm(P[:, :3], axis=1) - 1).max() < 1e-5
    pos, x, y, z = render.camera_frame(*render.CAMERAS["top"][:2])
    assert np.allclose(x, [1, 0, 0]) and np.allclose(y, [0, 1, 0]) and np.allclose(z, [0, 0, 1])      # straight above the table
    pos, x, y, z = render.camera_frame(*render.CAMERAS["angle"][:2])
    assert abs(x @ z) < 1e-12 and abs(y @ z) < 1e-12 and np.allclose(np.cross(x, y), z) and abs(x[2]) < 1e-12
    img, hit = render.render_numpy(m, np.concatenate([m["start_pose"][:6], [-0.2, 0.35, 0.02, 1, 0, 0, 0]]), 32, 24)
    assert img.shape == (24, 32, 3) and (hit == int(m["cg_cube"])).any() and (hit == int(m["cg_table"])).sum() > 200
