"""Golden vectors for the selective-scan recurrence from an INDEPENDENT, EXECUTED public implementation:
Hugging Face transformers' `MambaMixer.slow_forward` (transformers/models/mamba/modeling_mamba.py, the pure-PyTorch
path of state-spaces/mamba's selective scan -- the algorithm VMamba's `selective_scan_cuda_oflex`, the extension the
reference calls at Models/SS2D/csms6s.py:910,920-922, derives from).  The reference's own extension is not in
/root/reference and does not build offline, so this is the closest executable statement of the
operator's contract available in the build container.

The mixer is run in float64 on CPU; forward hooks capture the tensors that enter and leave its scan
(u = conv+act output, delta = dt_proj output incl. bias, B, C, A = -exp(A_log), D, gate) and autograd gives the
gradients at the same points.  Written to tests/golden/sscan_hf.npz:

    python tests/golden/make_golden_sscan_hf.py        (transformers 5.x, CPU)
"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ((2, 8, 1, 37), (1, 6, 4, 64), (2, 4, 16, 23))      # batch, channels (intermediate), d_state, L


def run_case(batch, dim, n, L, seed):
    from transformers.models.mamba.configuration_mamba import MambaConfig
    from transformers.models.mamba.modeling_mamba import MambaMixer

    torch.manual_seed(seed)
    cfg = MambaConfig(hidden_size=dim // 2, state_size=n, expand=2, conv_kernel=3, time_step_rank=3,
                      use_mambapy=False, num_hidden_layers=1, vocab_size=8)
    mixer = MambaMixer(cfg, layer_idx=0).double().train()
    assert mixer.intermediate_size == dim
    with torch.no_grad():                      # spread the parameters so that every term matters
        mixer.A_log.copy_(torch.log(0.3 + 2.0 * torch.rand(dim, n, dtype=torch.float64)))
        mixer.D.copy_(0.5 + torch.rand(dim, dtype=torch.float64))
        mixer.dt_proj.bias.copy_(torch.randn(dim, dtype=torch.float64) - 1.0)
        mixer.dt_proj.weight.mul_(3.0)
        mixer.x_proj.weight.mul_(4.0)
    cap = {}

    def keep(name):
        def hook(mod, inp, out):
            cap[name + "_in"] = inp[0]
            cap[name + "_out"] = out
            if out.requires_grad:
                out.retain_grad()
            if inp[0].requires_grad:
                inp[0].retain_grad()
        return hook

    mixer.x_proj.register_forward_hook(keep("x_proj"))       # in: u^T (b, L, dim); out: (b, L, R + 2n)
    mixer.dt_proj.register_forward_hook(keep("dt_proj"))     # out: delta (pre-softplus, bias included) (b, L, dim)
    mixer.in_proj.register_forward_hook(keep("in_proj"))     # out: (b, L, 2*dim) -> gate = second half
    mixer.out_proj.register_forward_hook(keep("out_proj"))   # in: (scan_output * act(gate))^T (b, L, dim)
    act_outs = []

    def keep_act(mod, inp, out):                              # 1st call: u = act(conv1d(.)), 2nd call: act(gate)
        out.retain_grad()
        act_outs.append(out)

    mixer.act.register_forward_hook(keep_act)
    x = torch.randn(batch, L, dim // 2, dtype=torch.float64)
    # slow_forward hard-codes `.float()` on A, B and u; keep those casts at float64 so that the fixture carries no
    # float32 input rounding (precision only -- the arithmetic is untouched)
    orig_float = torch.Tensor.float
    torch.Tensor.float = lambda self, *a, **k: self.double()
    try:
        y = mixer.slow_forward(x)
    finally:
        torch.Tensor.float = orig_float
    g_out = torch.randn_like(cap["out_proj_in"])              # upstream gradient at the gated scan output
    cap["out_proj_in"].backward(g_out)

    R = cfg.time_step_rank
    u = cap["x_proj_in"].transpose(1, 2)                      # (b, dim, L)
    assert act_outs[0].shape == u.shape and torch.equal(act_outs[0], u)
    ssm = cap["x_proj_out"]
    Bm = ssm[..., R:R + n].transpose(1, 2)                    # (b, n, L)
    Cm = ssm[..., R + n:].transpose(1, 2)
    delta = cap["dt_proj_out"].transpose(1, 2)                # (b, dim, L)
    gate = cap["in_proj_out"].transpose(1, 2)[:, dim:]        # (b, dim, L)
    act_gate = torch.nn.functional.silu(gate)
    out = {
        "u": u, "delta": delta, "A": -torch.exp(mixer.A_log), "B": Bm, "C": Cm, "D": mixer.D,
        "y_gated": cap["out_proj_in"].transpose(1, 2), "act_gate": act_gate,
        "dout": g_out.transpose(1, 2) * act_gate,             # gradient at the scan output itself
        "ddelta": cap["dt_proj_out"].grad.transpose(1, 2),
        "dB": ssm.grad[..., R:R + n].transpose(1, 2), "dC": ssm.grad[..., R + n:].transpose(1, 2),
        "dA": mixer.A_log.grad / (-torch.exp(mixer.A_log)),   # dL/dA from dL/dA_log (A = -exp(A_log))
        "dD": mixer.D.grad,
        "delta_bias": mixer.dt_proj.bias, "ddelta_bias": mixer.dt_proj.bias.grad,
        # total gradient at u (the act(conv1d) output) = the scan's du + the x_proj path (delta, B, C are functions
        # of u in the mixer); the test adds the latter to the oracle's du
        "du_total": act_outs[0].grad,
        "du_via_xproj": cap["x_proj_in"].grad.transpose(1, 2),
    }
    return {k: v.detach().numpy().copy() for k, v in out.items()}


def main():
    import transformers
    out = {"transformers_version": np.array(transformers.__version__)}
    for i, (b, d, n, L) in enumerate(CASES):
        for k, v in run_case(b, d, n, L, 100 + i).items():
            out[f"c{i}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "sscan_hf.npz"), **out)
    print("sscan_hf.npz", len(out), "arrays, transformers", transformers.__version__)


if __name__ == "__main__":
    main()
