"""Small end-to-end run that touches every kernel once (target of compute-sanitizer memcheck)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arshamg_scrnaseq_wgan_b200 import WGANConfig, WGANTrainer

dims = dict(noise_input_size=20, inflate_to_size=40, gex_size=133, disc_internal_size=24)
for two_cta in ("1", "0"):
    os.environ["WGAN_B200_GEMM_2CTA"] = two_cta
    B = 300
    tr = WGANTrainer(WGANConfig(n_critic=2, **dims), max_batch=B, device=0)
    tr.init_params(0)
    tr.set_data(torch.rand((500, dims["gex_size"]), device="cuda:0"))
    for it in range(2):
        out = tr.train_iteration_resident(it, B, read=True)
    x = torch.rand((77, dims["gex_size"]), device="cuda:0")
    z = tr.noise_prior(77, seed=1, call_id=3)
    tr.critic_step(x, z, tr.uniform(77, 1, 3))
    tr.generator_step(z)
    cells = tr.generator(z)
    d = tr.discriminator(cells)
    zfit, loss = tr.latent_fit(cells, steps=2)
    torch.cuda.synchronize()
    print("2cta", two_cta, out, float(d.mean()), float(loss.mean()))
    tr.close()
print("sanitize smoke done")
