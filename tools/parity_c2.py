"""BASELINE configs[1] parity run: flat ground, 4096 batched envs, random ctrl, fp64, 1000 steps on one B200; after
sampled steps every sampled env is re-stepped by the CPU oracle from the identical (qpos, qvel, act, qacc_warmstart,
ctrl) and compared (1e-9 relative on qpos / qvel / tendon length / reward-relevant COM velocity).
  python tools/parity_c2.py [--envs 4096] [--steps 1000] [--every 10] [--sample 512] [--xml flat] [--procs 16]
Prints one JSON line."""
import argparse, json, multiprocessing as mp, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

_mj = None


def _init(xml):
    global _mj
    from oracle import oracle as O
    _mj = O.MjLike(xml)


def _check(job):
    qpos, qvel, act, warm, ctrl, qpos1, qvel1, ten1, nmpr = job
    mj = _mj
    mj.reset_data()
    mj.qpos[:] = qpos; mj.qvel[:] = qvel; mj.act[:] = act; mj.qacc_warmstart[:] = warm; mj.ctrl[:] = ctrl
    mj.step(20)
    rel = lambda a, b: float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))
    return rel(qpos1, mj.qpos), rel(qvel1, mj.qvel), rel(ten1, mj.ten_length), float(nmpr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--every", type=int, default=10)
    ap.add_argument("--sample", type=int, default=512)
    ap.add_argument("--xml", default="flat")
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--dump", default="", help="npz file receiving the inputs / GPU outputs of the outlier cases")
    a = ap.parse_args()
    pool = mp.get_context("fork").Pool(a.procs, initializer=_init, initargs=(a.xml,))   # fork before CUDA is touched
    import torch
    from tensegrity_rl_b200 import TensegrityVecEnv
    n = a.envs
    v = TensegrityVecEnv(n, xml_file=a.xml, env="tr_env", auto_reset=False, terminate_when_unhealthy=False, max_episode_steps=0)
    v.reset_tensor()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    rng = np.random.default_rng(0)
    worst = np.zeros(3); nbad = ncheck = 0; out_max = 0.0; out_with_mpr = 0; n_with_mpr = 0; hist = np.zeros(8, int); t0 = time.time()
    pending = []
    for step in range(a.steps):
        act = -0.45 + 0.3 * torch.rand(n, 6, generator=g, device="cuda", dtype=torch.float64)
        check = step % a.every == a.every - 1
        if check:
            before = v.get_state()
        v.step_tensor(act)
        if not check:
            continue
        after, info = v.get_state(), v.info.cpu().numpy()
        idx = rng.choice(n, min(a.sample, n), replace=False)
        jobs = [(before["qpos"][e], before["qvel"][e], before["act"][e], before["qacc_warmstart"][e], after["ctrl"][e],
                 after["qpos"][e], after["qvel"][e], info[e, 8:17], info[e, 30]) for e in idx]
        pending.append((pool.map_async(_check, jobs, chunksize=16), jobs, info[idx]))
    dumps = []
    for p, jobs, rows in pending:
        for r, job, row in zip(p.get(), jobs, rows):
            m = max(r[:3]); ncheck += 1; n_with_mpr += r[3] > 0
            hist[min(7, max(0, int(np.floor(np.log10(max(m, 1e-16))) + 16)))] += 1
            if m > 1e-9:
                nbad += 1; out_max = max(out_max, m); out_with_mpr += r[3] > 0
                dumps.append(np.concatenate([np.ravel(x) for x in job[:8]] + [row, [m]]))
            else:
                worst = np.maximum(worst, r[:3])
    info = v.info.cpu().numpy()
    print(json.dumps({"config": "BASELINE configs[1]: %s XML, tr_env, %d envs, random ctrl U[-0.45,-0.15], fp64, %d steps" % (a.xml, n, a.steps),
                      "checked_env_steps": ncheck, "outliers_above_1e-9": nbad, "largest_outlier": out_max,
                      "outliers_in_steps_that_ran_MPR": int(out_with_mpr), "checked_env_steps_that_ran_MPR": int(n_with_mpr),
                      "worst_rel_within_tol": {"qpos": worst[0], "qvel": worst[1], "ten_length": worst[2]},
                      "log10_error_histogram_1e-16_to_1e-9+": hist.tolist(), "tolerance": 1e-9,
                      "final_step_mean_contacts": float(info[:, 19].mean()), "contact_overflow": int(info[:, 28].sum()),
                      "bad_state": int(info[:, 29].sum()), "seconds": time.time() - t0}))
    if a.dump and dumps:
        np.savez(a.dump, cases=np.array(dumps))   # qpos 21, qvel 18, act 6, warm 18, ctrl 6, qpos' 21, qvel' 18, ten' 9, info 32, err
    pool.close()


if __name__ == "__main__":
    main()
