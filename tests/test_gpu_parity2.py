"""GPU parity tests added in round 2 (``pytest -m gpu``): the size holes the round-1 review listed -- fallback
rows of the hull fixture, irregular R = 8 at the benchmark's batch size, the 32768-bin allocation, buffered k = 10
-- plus the on-device item generator, sequence reloads and (when two GPUs are visible) handles on two devices.
Same bar as tests/test_gpu_parity.py: value-exact through the C ABI against the oracle / the committed outputs
of the unmodified reference."""
import numpy as np
import pytest

from conftest import lib_from_fixture, load_golden

pytestmark = pytest.mark.gpu


def _env(lib, seqs, **kw):
    from irbpp_b200.vec_env import GpuVecEnv
    return GpuVecEnv(lib, seqs, device=kw.pop("device", "cuda:0"), **kw)


def test_hull_fixture_including_no_candidate_fallback_rows():
    """Every map of the hull fixture through select / pad (binPhy.py:205-225): where the verbatim reference found
    no hull candidate (counts == 0) the table is the fallback -- the selectedAction smallest posZValid in stable
    order, H = bin height, V = naiveMask -- compared row by row with the oracle's select_candidates."""
    from irbpp_b200 import shapes
    from oracle.oracle_env import OracleConfig, select_candidates
    d = load_golden("hulls")
    n = len(d["counts"])
    lib = shapes.make_cube_library(seed=1, num_rotations=1, num_shapes=4)
    sel = 256
    env = _env(lib, np.zeros((n, 8), np.int32), selected_action=sel)
    out = env.debug_hulls(d["posZValid"][:, None], d["mask"][:, None])
    cfg = OracleConfig(ZRotNum=1, selectedAction=sel)
    off, n_fallback = 0, 0
    for k in range(n):
        c = int(d["counts"][k])
        rows = d["rows"][off:off + c] if c else None
        want = select_candidates(cfg, rows, d["posZValid"][k][None], d["mask"][k][None])
        assert np.array_equal(out["cand"][k], want), k
        n_fallback += (c == 0)
        off += c
    assert n_fallback > 0
    # an all-infeasible and a barely-feasible map: both fallback branches (mask empty / mask set but no hull is impossible
    # for a non-empty mask, so the second one checks that a single feasible pose is a hull candidate)
    pv = np.full((n, 1, 16, 16), 1e3); mk = np.zeros((n, 1, 16, 16))
    pv[1, 0, 3, 4] = 0.05; mk[1, 0, 3, 4] = 1.0
    out = env.debug_hulls(pv, mk)
    want0 = select_candidates(cfg, None, pv[0], mk[0])
    assert out["num_hull"][0] == 0 and np.array_equal(out["cand"][0], want0)
    assert out["num_hull"][1] == 1 and np.array_equal(out["cand"][1][0], [0, 3, 4, 0.05, 1.0])
    env.close()


@pytest.mark.parametrize("kind,R,n_sub,steps", [("irregular", 8, 24, 30), ("cube", 2, 24, 40)])
def test_full_size_general_subset_matches_oracle(kind, R, n_sub, steps):
    """4096 bins of the general (irregular, R = 8: the dense-pose scan, truncation to 500 rows) and Cube workloads
    on the GPU; a random subset replayed by the oracle with the same actions, value-exact."""
    import torch
    from irbpp_b200 import shapes
    from oracle.oracle_env import OracleConfig, OracleEnv
    n = 4096
    lib = shapes.make_irregular_library(32, seed=2, num_rotations=8) if kind == "irregular" else shapes.make_cube_library(seed=3)
    seqs = shapes.make_sequences(n, 128, lib.num_shapes, seed=0)
    env = _env(lib, seqs)
    rng = np.random.default_rng(12)
    subset = np.sort(rng.choice(n, size=n_sub, replace=False))
    cfg = OracleConfig(ZRotNum=R)
    oracles = [OracleEnv(cfg, lib, seqs[i]) for i in subset]
    obs = env.reset()
    want = np.stack([o.reset() for o in oracles])
    assert np.array_equal(obs[subset].cpu().numpy(), want.astype(np.float32))
    gen = torch.Generator(device="cuda:0"); gen.manual_seed(4)
    n_done = 0
    for t in range(steps):
        mask = obs[:, :2500].view(n, 500, 5)[:, :, 4] == 1
        acts = torch.argmax(torch.rand((n, 500), device="cuda:0", generator=gen) + mask.float(), dim=1)
        obs, rew, done, infos = env.step(acts)
        a = acts.cpu().numpy()
        for k, i in enumerate(subset):
            o, r, dn, info = oracles[k].step(int(a[i]))
            if dn:
                assert infos[int(i)]["counter"] == info["counter"] and infos[int(i)]["ratio"] == info["ratio"]
                o = oracles[k].reset()
                n_done += 1
            want[k] = o
            assert np.float32(r) == rew[int(i), 0].item() and bool(dn) == bool(done[i])
        assert np.array_equal(obs[subset].cpu().numpy(), want.astype(np.float32)), (kind, t)
    assert n_done > 0
    env.close()


def test_32768_bins_on_one_gpu_subset_matches_oracle():
    """BASELINE.json configs[4] holds 32768 bins; here all of them live on ONE GPU (8 x the per-GPU share of the
    8-GPU run): allocation, indexing beyond 2^15 bins, a subset against the oracle."""
    import torch
    from irbpp_b200 import shapes
    from oracle.oracle_env import OracleConfig, OracleEnv
    n = 32768
    lib = shapes.make_irregular_library(32, seed=2, num_rotations=8)
    seqs = shapes.make_sequences(n, 32, lib.num_shapes, seed=0)
    env = _env(lib, seqs)
    subset = np.array([0, 1, 4095, 4096, 16383, 20000, 32766, 32767])
    cfg = OracleConfig(ZRotNum=8)
    oracles = [OracleEnv(cfg, lib, seqs[i]) for i in subset]
    obs = env.reset()
    want = np.stack([o.reset() for o in oracles])
    assert np.array_equal(obs[subset].cpu().numpy(), want.astype(np.float32))
    gen = torch.Generator(device="cuda:0"); gen.manual_seed(5)
    for t in range(12):
        mask = obs[:, :2500].view(n, 500, 5)[:, :, 4] == 1
        acts = torch.argmax(torch.rand((n, 500), device="cuda:0", generator=gen) + mask.float(), dim=1)
        obs, rew, done, infos = env.step(acts)
        a = acts.cpu().numpy()
        for k, i in enumerate(subset):
            o, r, dn, info = oracles[k].step(int(a[i]))
            if dn:
                o = oracles[k].reset()
            want[k] = o
            assert bool(dn) == bool(done[i])
        assert np.array_equal(obs[subset].cpu().numpy(), want.astype(np.float32)), t
    env.close()


def test_buffered_k10_episode_matches_reference_golden():
    """BASELINE.json configs[3]: buffered k = 10.  Trace generated by tests/golden/make_golden.py from the
    UNMODIFIED reference geometry (episode_buffered10.npz): order observation, get_action_candidates, step."""
    d = load_golden("episode_buffered10")
    lib = lib_from_fixture(d)
    assert int(d["bufferSize"]) == 10
    env = _env(lib, d["sequences"], buffer_size=10)
    obs = env.reset()
    assert np.array_equal(obs.cpu().numpy(), d["obs"][0].astype(np.float32))
    for t in range(len(d["actions"])):
        loc = env.get_action_candidates(d["order"][t], as_tensor=True)
        assert np.array_equal(loc.cpu().numpy(), d["loc_obs"][t].astype(np.float32)), t
        obs, rew, done, infos = env.step(d["actions"][t])
        assert np.array_equal(obs.cpu().numpy(), d["obs"][t + 1].astype(np.float32)), t
        assert np.array_equal(rew.numpy()[:, 0], d["reward"][t].astype(np.float32))
        assert np.array_equal(done, d["done"][t])
    env.close()


def test_all_possible_observation_k10_matches_oracle():
    """get_all_possible_observation at k = 10 (one fused pass over all buffer slots) against the oracle."""
    from irbpp_b200 import shapes
    from oracle.oracle_env import OracleConfig, OracleVecEnv
    lib = shapes.make_blockout_library(16, seed=41)
    n, k = 12, 10
    seqs = shapes.make_sequences(n, 64, lib.num_shapes, seed=9)
    cfg = OracleConfig(ZRotNum=4, bufferSize=k)
    ora = OracleVecEnv(cfg, lib, seqs)
    env = _env(lib, seqs, buffer_size=k)
    assert np.array_equal(env.reset().cpu().numpy(), ora.reset().astype(np.float32))
    rng = np.random.default_rng(2)
    for t in range(8):
        want = np.stack([e.get_all_possible_observation() for e in ora.envs])
        got = env.get_all_possible_observation()
        assert np.array_equal(got.cpu().numpy(), want.astype(np.float32)), t
        order = rng.integers(0, k, size=n)
        loc_o = np.stack(ora.get_action_candidates(order))
        loc_g = env.get_action_candidates(order, as_tensor=True)
        assert np.array_equal(loc_g.cpu().numpy(), loc_o.astype(np.float32)), t
        acts = np.zeros(n, dtype=np.int64)
        for i in range(n):
            valid = np.nonzero(loc_o[i][:2500].reshape(500, 5)[:, 4] == 1)[0]
            acts[i] = int(rng.choice(valid)) if len(valid) else 0
        o_obs = ora.step(acts)[0]
        g_obs = env.step(acts)[0]
        assert np.array_equal(g_obs.cpu().numpy(), o_obs.astype(np.float32)), t
    env.close()


def test_item_generator_matches_oracle_stream():
    """sequences=None: ids drawn on the device (irbpp_set_item_rng).  The oracle is given the same stream
    (shapes.item_rng_ids, the Python mirror of the device mixer) as an explicit, long-enough sequence."""
    from irbpp_b200 import shapes
    from oracle.oracle_env import OracleConfig, OracleVecEnv
    lib = shapes.make_blockout_library(24, seed=31)
    n, steps = 40, 50
    stream = shapes.item_rng_ids(20240924, n, 4 * steps + 8, lib.num_shapes)
    ora = OracleVecEnv(OracleConfig(ZRotNum=4), lib, stream)
    env = _env(lib, None, num_envs=n, item_seed=20240924)
    o_obs = ora.reset()
    assert np.array_equal(env.reset().cpu().numpy(), o_obs.astype(np.float32))
    rng = np.random.default_rng(6)
    for t in range(steps):
        acts = np.zeros(n, dtype=np.int64)
        for i in range(n):
            valid = np.nonzero(o_obs[i][:2500].reshape(500, 5)[:, 4] == 1)[0]
            acts[i] = int(rng.choice(valid)) if len(valid) else 0
        o_obs, o_rew, o_done, _ = ora.step(acts)
        g_obs, g_rew, g_done, _ = env.step(acts)
        assert np.array_equal(g_obs.cpu().numpy(), o_obs.astype(np.float32)), t
        assert np.array_equal(g_done, o_done)
    assert max(e.cursor for e in ora.envs) < stream.shape[1]          # the oracle never wrapped its copy of the stream
    assert np.array_equal(env.debug_state()["cursor"], [e.cursor for e in ora.envs])
    env.close()


def test_sequences_can_be_reloaded():
    """irbpp_set_sequences twice on one handle: the second call replaces the pool and restarts the cursors."""
    from irbpp_b200 import shapes
    from oracle.oracle_env import OracleConfig, OracleVecEnv
    lib = shapes.make_blockout_library(16, seed=3)
    n = 8
    s1 = shapes.make_sequences(n, 5, lib.num_shapes, seed=1)          # short: wraps within the test
    s2 = shapes.make_sequences(n, 7, lib.num_shapes, seed=2)
    env = _env(lib, s1)
    for seqs in (s1, s2):
        if seqs is s2:
            env._check(env._lib.irbpp_set_sequences(env._h, seqs.ctypes.data, seqs.shape[1]))
        ora = OracleVecEnv(OracleConfig(ZRotNum=4), lib, seqs)
        assert np.array_equal(env.reset().cpu().numpy(), ora.reset().astype(np.float32))
        for t in range(14):
            acts = np.zeros(n, dtype=np.int64)
            assert np.array_equal(env.step(acts)[0].cpu().numpy(), ora.step(acts)[0].astype(np.float32)), t
        assert np.array_equal(env.debug_state()["cursor"], [e.cursor for e in ora.envs])
    env.close()


def test_handles_on_two_devices():
    """cudaFuncAttributeMaxDynamicSharedMemorySize is per device: a handle on cuda:1 created after one on cuda:0
    must get its own opt-in (irregular tables need > 48 KB of dynamic shared memory in the scan kernel)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    from irbpp_b200 import shapes
    from oracle.oracle_env import OracleConfig, OracleVecEnv
    lib = shapes.make_irregular_library(8, seed=5, num_rotations=24)
    seqs = shapes.make_sequences(6, 16, lib.num_shapes, seed=1)
    envs = [_env(lib, seqs, device="cuda:%d" % d) for d in (0, 1)]
    ora = OracleVecEnv(OracleConfig(ZRotNum=24), lib, seqs)
    want = ora.reset().astype(np.float32)
    for e in envs:
        assert np.array_equal(e.reset().cpu().numpy(), want)
    acts = np.zeros(6, dtype=np.int64)
    for t in range(4):
        want = ora.step(acts)[0].astype(np.float32)
        for e in envs:
            assert np.array_equal(e.step(acts)[0].cpu().numpy(), want), t
    for e in envs:
        e.close()
    assert torch.cuda.current_device() == 0          # the library leaves the caller's current device alone


def test_point_cloud_sampling_and_shape_features_match_torch_fp32():
    """SURVEY.md 8(f)3 on the GPU: DeviceShapeClouds.sample == shapeArray[ids][:, indices] (exact) and
    DeviceShapeClouds.features == max(shapeEncoder(nextShape)) of model.py:328-335 evaluated by torch in float32 on
    the same device.  Tolerance (stated): |diff| <= 2e-5 * (1 + |want|) -- the fused kernel accumulates the 128
    products of the second layer with FMA in index order, cuBLAS in its own order (possibly TF32-free fp32)."""
    import torch
    from irbpp_b200.pointnet import DeviceShapeClouds, pn_indices
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(5)
    S, Pn, B = 32, 20000, 4096
    shape_array = (torch.rand((S, Pn, 3), generator=g) * 0.15).float()
    ids = torch.randint(0, S, (B,), generator=g)
    obs = torch.zeros((B, 3533)); obs[:, 2500] = ids.float()
    obs = obs.to("cuda:0")
    enc = torch.nn.Sequential(torch.nn.Linear(3, 128), torch.nn.LeakyReLU(), torch.nn.Linear(128, 128), torch.nn.LeakyReLU()).cuda()
    clouds = DeviceShapeClouds(shape_array, device="cuda:0", n_points=1024, seed=11)
    for counter in (0, 1, 7):
        idx = torch.from_numpy(pn_indices(11, counter, 1024, Pn))
        want_cloud = shape_array[ids][:, idx].cuda()                      # the reference's host gather + H2D
        got_cloud, got_idx = clouds.sample(obs, counter=counter, return_indices=True)
        assert torch.equal(got_idx.cpu().long(), idx) and torch.equal(got_cloud, want_cloud)
        with torch.no_grad():
            want = torch.max(enc(want_cloud), dim=1)[0]
        got = clouds.features(obs, enc, counter=counter)
        assert torch.all((got - want).abs() <= 2e-5 * (1 + want.abs())), float((got - want).abs().max())
        got2 = clouds.features(ids.cuda(), enc, counter=counter)         # explicit ids instead of the observation column
        assert torch.equal(got, got2)


def test_actor_loop_through_learner_glue_at_4096_bins():
    """SURVEY.md 8(f)2, executed: the reference's actor loop shape (trainer.py:157-186) for 60 iterations at N = 4096
    -- mask on the device (tools.py:283-300), infos consumed through the batched views, one replay-bank append per
    step, a batch drawn with the segment rule of agent.py:69 made safe for N > batch_size -- and the batched views
    agree with the per-bin dicts the reference's loop reads."""
    import torch
    from irbpp_b200 import shapes, learner_glue as glue
    n, sel = 4096, 500
    lib = shapes.make_blockout_library(32, seed=1)
    env = _env(lib, None, num_envs=n, item_seed=3)
    gen = torch.Generator(device="cuda:0"); gen.manual_seed(1)
    bank = glue.ReplayBank(n, 8, env.obs_len, "cuda:0")
    stats = glue.EpisodeStats()
    state = env.reset()
    assert glue.segment_size(64, 4096) == (64, 1) and glue.segment_size(64, 16) == (16, 4)
    total_done = 0
    for T in range(1, 61):
        mask = glue.get_mask_from_state(state, sel)
        assert mask.shape == (n, sel) and mask.is_cuda
        q = torch.rand((n, sel), device="cuda:0", generator=gen)
        q[(1 - mask).bool()] = -float("inf")
        action = q.argmax(1)
        next_state, reward, done, infos = env.step(action.cpu().numpy())
        valid = stats.update(done, infos)
        idx, ep_r, ratio, counter, v = infos.finished()
        assert np.array_equal(idx, np.nonzero(done)[0]) and valid.all()
        for j in idx[:5]:                                   # the batched views say what the per-bin dicts say
            info = infos[int(j)]
            k = int(np.nonzero(idx == j)[0][0])
            assert info["episode"]["r"] == ep_r[k] and info["ratio"] == ratio[k] and info["counter"] == counter[k]
        total_done += len(idx)
        dv = env.last_step_device()                          # device copies of the step's results == what step() returned
        assert np.array_equal(dv["reward"].cpu().numpy(), reward.numpy()[:, 0]) and np.array_equal(dv["done"].cpu().numpy() != 0, done)
        if T % 2:
            bank.append_batch(state, action, reward, done, valid)
        else:
            bank.append_from_env(env, state, action)        # the same transition without the host round trip
            slot = (bank.t - 1) % bank.cap
            assert torch.equal(bank.rewards[slot].cpu(), reward[:, 0]) and np.array_equal(bank.nonterminal[slot].cpu().numpy(), ~done)
        if T % 4 == 0:
            envs, slots, s, a, r, s2, nonterm = bank.sample(64, generator=gen)
            assert s.shape == (64, env.obs_len) and s2.shape == s.shape and a.shape == (64,)
            assert torch.equal(s, bank.states[slots, envs])
        state = next_state
    assert stats.episodes == total_done and total_done > 0 and len(stats.episode_ratio) > 0
    env.close()


def test_single_env_evaluation_surface_matches_oracle():
    """The loop of tools.test (tools.py:303-358): reset / step(int) / get_ratio() before the explicit reset() /
    .packed / .item_creator.traj_index on a single environment, against the (non-vectorised) oracle env driven the
    same way -- including that the reset() after a finished episode does not draw a second item."""
    from irbpp_b200 import shapes
    from irbpp_b200.single_env import SingleGpuEnv
    from oracle.oracle_env import OracleConfig, OracleEnv
    lib = shapes.make_blockout_library(16, seed=7)
    seq = shapes.make_sequences(1, 200, lib.num_shapes, seed=4)[0]
    env = SingleGpuEnv(lib, seq, device="cuda:0")
    ora = OracleEnv(OracleConfig(ZRotNum=4), lib, seq)
    rng = np.random.default_rng(8)
    done, episodes = True, 0
    for _ in range(120):
        if done:
            state, o_state = env.reset(), ora.reset()
            done = False
        assert np.array_equal(state, o_state.astype(np.float32))
        valid = np.nonzero(o_state[:2500].reshape(500, 5)[:, 4] == 1)[0]
        a = int(rng.choice(valid)) if len(valid) and rng.random() > 0.05 else int(rng.integers(0, 500))
        state, reward, done, info = env.step(a)
        o_state, o_reward, o_done, o_info = ora.step(a)
        assert np.float32(o_reward) == np.float32(reward) and done == o_done
        if done:
            assert env.get_ratio() == ora.get_ratio() == info["ratio"] and info["counter"] == o_info["counter"] == len(ora.packed_ids)
            assert [p[0] for p in env.packed] == ora.packed_ids
            episodes += 1
        else:
            assert env.get_ratio() == ora.get_ratio()
    assert episodes >= 2 and env.item_creator.traj_index == episodes + (0 if done else 1)
    env.close()
