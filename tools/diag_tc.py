"""Stage-by-stage comparison of the CUDA path with the dense oracle (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
from helpers import build_dropin, to_dev, normwise
from raindrop_b200 import functional as RF, lib as L
from raindrop_b200.synth import model_config, make_batch, synth_weights, used_param_keys
from oracle.raindrop_oracle import build_oracle_model

def run(cfg_name, B, seed):
    cfg = model_config(cfg_name, dropout=0.2)
    batch = make_batch(cfg, B, seed=seed)
    oracle = build_oracle_model(cfg).eval(); synth_weights(oracle, cfg, seed=21)
    st = {}
    ref, _, _ = oracle.forward_dense(batch["src"], batch["static"], batch["times"], batch["lengths"], stages=st)
    F.cross_entropy(ref, batch["y"]).backward()
    go = dict(oracle.named_parameters())
    for rep in range(2):
        model = build_dropin(cfg, 21).eval()
        d = to_dev(batch)
        logits, _, _ = model.forward(d["src"], d["static"], d["times"], d["lengths"])
        F.cross_entropy(logits, d["y"]).backward()
        N, C = cfg["d_inp"], cfg["max_len"] * 4
        x0 = RF.workspace_view(model._plan, L.WS_X0).view(B * N, C).cpu()
        h1 = RF.workspace_view(model._plan, L.WS_H1).view(B * N, C).cpu()
        T = cfg["max_len"]; D = N * 4 + 16
        enc_in = RF.workspace_view(model._plan, L.WS_ENC_IN).view(T, B, D).cpu()
        obs = enc_in[:, :, :N * 4]
        e_x0 = normwise(x0, st["x0"].reshape(B * N, C)); e_h1 = normwise(h1, st["h1"].reshape(B * N, C).detach())
        e_obs = normwise(obs, st["obs"].detach())
        # where is the obs error?
        diff = (obs - st["obs"].detach()).abs()
        idx = torch.nonzero(diff > 1e-3 * st["obs"].abs().max())
        gp = dict(model.named_parameters())
        ge = {k.split(".lin_value.")[0][-6:] + "." + k.split(".")[-1]: normwise(gp[k].grad, go[k].grad)
              for k in used_param_keys(cfg) if "lin_value" in k}
        print("%s B=%d rep%d: x0 %.2e h1 %.2e obs %.2e logits %.2e | bad obs elems %d %s | grads %s" % (
            cfg_name, B, rep, e_x0, e_h1, e_obs, normwise(logits, ref), idx.shape[0],
            idx[:6].tolist(), {k: "%.1e" % v for k, v in ge.items()}))
        worst = max(((k, normwise(gp[k].grad, go[k].grad)) for k in used_param_keys(cfg)), key=lambda kv: kv[1])
        print("    worst grad:", worst)

for name, B, seed in (("P19", 4, 15), ("P19", 5, 16), ("P19", 37, 137), ("P12", 2, 17), ("PAM", 2, 18), ("TINY", 3, 11)):
    run(name, B, seed)
