"""Host-mapped I/O of the step (include/dmb.h "Host-mapped I/O", BatchedSim.step_host): the kernel reads the action
rows from, and writes the record rows to, pinned host memory.  The result must be bit-identical to the device-buffer
path (same kernel, only the addresses differ), including auto-resets, and the N = 1 gym surface -- which steps through
it -- must agree with a batched env of one."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_step_host_bit_identical_to_device_path():
    from deepmimic_mujoco_b200.env import DPVecEnv
    n = 300
    kw = dict(motions=("walk", "spinkick"), seed=11, reward_mode=4, auto_reset=True,
              clip_ids=torch.arange(n, dtype=torch.int32) % 2)
    a, b = DPVecEnv(n, **kw), DPVecEnv(n, **kw)
    assert torch.equal(a.reset(), b.reset())
    hact, hrec = b.sim.enable_host_io()
    hact2 = b.sim.alloc_host((n, 28))                       # caller-rotated action buffers work as well
    rng = np.random.default_rng(0)
    ndone = 0
    for t in range(40):
        act = (rng.random((n, 28), dtype=np.float32) - 0.5) * (3.0 if t % 7 == 3 else 1.0)
        obs, rew, done, info = a.step(torch.from_numpy(act).cuda())
        dact = torch.from_numpy(act).cuda()
        if t % 3 == 0:                                       # both sides host-mapped
            hact.array[:] = act
            assert b.step_host(hact, hrec) is hrec
        elif t % 3 == 1:                                     # device action, record straight to the host
            b.step(dact, rec_host=hrec)
        else:                                                # host-mapped action (a caller-rotated buffer), device record
            hact2.array[:] = act
            b.step(hact2)
        torch.cuda.synchronize()
        rec = a.sim.rec.cpu().numpy()
        got = hrec.array if t % 3 != 2 else b.sim.rec.cpu().numpy()
        assert np.array_equal(got, rec), t
        assert torch.equal(b.sim.obs, obs) and torch.equal(b.sim.reward, rew) and torch.equal(b.sim.done, done)
        assert torch.equal(b.sim.last_len, info["episode_length"])
        ndone += int(done.sum())
    assert ndone > 0                                         # auto-resets happened on the way
    for k in ("qpos", "qvel", "warm", "idx_curr", "ep_len", "reset_count"):
        assert torch.equal(getattr(a.sim, k), getattr(b.sim, k)), k
    with pytest.raises(ValueError):
        b.step_host(hrec, hrec)                              # wrong shape for the action array
    a.close(); b.close()
    assert hact.dev_ptr == 0 and hrec.array is None          # freed with the sim


def test_gym_surface_steps_through_host_io_and_matches_batched_env():
    from deepmimic_mujoco_b200.env import DPEnv, DPVecEnv
    env = DPEnv(motion="walk", seed=0)
    vec = DPVecEnv(1, motions=("walk",), seed=0, auto_reset=False)
    env.reset()
    q, v = env.qpos, env.qvel
    env.set_state(q, v)                                      # leaves warmstart = qacc of mj_forward(ctrl = 0)
    vec.sim.set_state(q[None], v[None], warm=env._sim.warm[:, : env._sim.nv])
    vec.sim.idx_curr.copy_(env._sim.idx_curr); vec.sim.idx_init.copy_(env._sim.idx_init)
    launches0 = env._sim.kernel_launches()
    rng = np.random.default_rng(1)
    for t in range(12):
        ac = rng.uniform(-0.5, 0.5, 28).astype(np.float32)
        ob, rew, done, _ = env.step(ac)
        o2, r2, d2, _ = vec.step(torch.from_numpy(ac[None]).cuda())
        assert ob.dtype == np.float64 and np.array_equal(ob, o2[0].double().cpu().numpy()), t
        assert rew == float(r2[0]) and done == bool(d2[0])
        assert np.array_equal(env.sim.data.ctrl, ac.astype(np.float64))
        if done:
            break
    # one dmb_step per env.step: the step kernel (+ the scheduler sort, every 8th step when one round holds all envs)
    steps = t + 1
    assert steps <= env._sim.kernel_launches() - launches0 <= 2 * steps
    env.close(); vec.close()


def test_device_pointer_of_torch_pinned_memory():
    """A buffer the caller pinned itself (torch pin_memory) can be passed after dmb_host_device_pointer resolved it."""
    from deepmimic_mujoco_b200 import lib as _lib
    from deepmimic_mujoco_b200.env import DPVecEnv
    L = _lib.load()
    n = 64
    env, ref = DPVecEnv(n, seed=2), DPVecEnv(n, seed=2)
    env.reset(); ref.reset()
    act = (torch.rand(n, 28) - 0.5).pin_memory()
    dp = C.c_void_p()
    rc = L.dmb_host_device_pointer(env.sim.device.index, C.c_void_p(act.data_ptr()), C.byref(dp))
    if rc != 0:
        pytest.skip("torch's pinned allocator does not map its memory into the device address space on this box")
    assert dp.value
    s = env.sim
    with torch.cuda.device(s.device):
        _lib.check(L.dmb_step(s.handle, C.byref(s._st), dp, C.byref(s._out), s._stream()), s.handle, "dmb_step")
    ref.step(act.cuda())
    torch.cuda.synchronize()
    assert torch.equal(s.obs, ref.sim.obs) and torch.equal(s.reward, ref.sim.reward)
    env.close(); ref.close()
