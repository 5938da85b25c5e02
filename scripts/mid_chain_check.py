"""Fused row-local chains vs the separate GEMMs + kernels (WGAN_B200_NO_MID_FUSE=1): same gradients / scalars."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer
    out = {}
    for form in ("explicit", "gram"):
        for dims, B in ((dict(), 4096), (dict(noise_input_size=20, inflate_to_size=40, gex_size=133, disc_internal_size=24), 77),
                        (dict(), 200)):
            tr = WGANTrainer(WGANConfig(gp_form=form, **dims), max_batch=B, device=0)
            tr.init_params(seed=1)
            g = torch.Generator(device="cuda:0").manual_seed(0)
            x = torch.rand((B, tr.cfg.gex_size), device="cuda:0", generator=g)
            z = tr.noise_prior(B, seed=5, call_id=1, row0=0)
            eps = tr.uniform(B, seed=5, call_id=1, row0=0)
            sc = tr.critic_step(x, z, eps, apply=False)
            gc = tr.arena(1, 1).clone().cpu().numpy()
            sg = tr.generator_step(z, apply=False)
            gg = tr.arena(0, 1).clone().cpu().numpy()
            key = f"{form}-{B}"
            np.save(f"/tmp/mid_{os.environ.get('WGAN_B200_NO_MID_FUSE','0')}_{key}_c.npy", gc)
            np.save(f"/tmp/mid_{os.environ.get('WGAN_B200_NO_MID_FUSE','0')}_{key}_g.npy", gg)
            out[key] = [sc, sg]
            tr.close()
    print(json.dumps(out))
else:
    import numpy as np
    res = {}
    for v in ("0", "1"):
        env = dict(os.environ)
        if v == "1": env["WGAN_B200_NO_MID_FUSE"] = "1"
        else: env.pop("WGAN_B200_NO_MID_FUSE", None)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=300)
        if r.returncode != 0:
            print("child failed", v, r.stderr[-2000:]); sys.exit(1)
        res[v] = json.loads(r.stdout.strip().splitlines()[-1])
    for key in res["0"]:
        a, b = res["0"][key], res["1"][key]
        gc0, gc1 = np.load(f"/tmp/mid_0_{key}_c.npy"), np.load(f"/tmp/mid_1_{key}_c.npy")
        gg0, gg1 = np.load(f"/tmp/mid_0_{key}_g.npy"), np.load(f"/tmp/mid_1_{key}_g.npy")
        fc = np.linalg.norm(gc0 - gc1) / np.linalg.norm(gc1)
        fg = np.linalg.norm(gg0 - gg1) / np.linalg.norm(gg1)
        print(key, "critic grad fro", f"{fc:.2e}", "gen grad fro", f"{fg:.2e}", "scalars fused", a, "unfused", b)
