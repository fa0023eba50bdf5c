"""CPU baseline: the reference's generate() loop restated with the same torch CPU library calls
(nn.Linear / nn.GRUCell / F.softmax, fatchord_version.py:155-236) so that its speed is representative
of the reference's PyTorch CPU path on the box's host cores.

TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline and --impl reference legs).  The unmodified
reference cannot travel to the GPU box (it is a Python tree outside this repo), hence kind = "port".
Checked against oracle/wavernn_oracle.py in tests/test_oracle_golden.py::test_torch_port_matches_oracle.
"""
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import philox
from . import wavernn_oracle as orc
from .weights import AUX_DIMS, RNN_DIMS


class TorchPort:
    def __init__(self, sd, mode="RAW", threads=None):
        if threads:
            torch.set_num_threads(int(threads))
        self.mode = mode
        self.sd = sd
        t = lambda k: torch.from_numpy(np.ascontiguousarray(sd[k]))
        self.I = torch.nn.Linear(112, RNN_DIMS)
        self.rnn1 = torch.nn.GRUCell(RNN_DIMS, RNN_DIMS)
        self.rnn2 = torch.nn.GRUCell(RNN_DIMS + AUX_DIMS, RNN_DIMS)
        self.fc1 = torch.nn.Linear(RNN_DIMS + AUX_DIMS, RNN_DIMS)
        self.fc2 = torch.nn.Linear(RNN_DIMS + AUX_DIMS, RNN_DIMS)
        self.C = sd["fc3.weight"].shape[0]
        self.fc3 = torch.nn.Linear(RNN_DIMS, self.C)
        with torch.no_grad():
            self.I.weight.copy_(t("I.weight")); self.I.bias.copy_(t("I.bias"))
            for cell, n in ((self.rnn1, "rnn1"), (self.rnn2, "rnn2")):      # get_gru_cell, :267-273
                cell.weight_ih.copy_(t(n + ".weight_ih_l0")); cell.weight_hh.copy_(t(n + ".weight_hh_l0"))
                cell.bias_ih.copy_(t(n + ".bias_ih_l0")); cell.bias_hh.copy_(t(n + ".bias_hh_l0"))
            for lin, n in ((self.fc1, "fc1"), (self.fc2, "fc2"), (self.fc3, "fc3")):
                lin.weight.copy_(t(n + ".weight")); lin.bias.copy_(t(n + ".bias"))

    @torch.no_grad()
    def loop(self, mels, aux, seed=0, max_steps=None, time_budget_s=None):
        """mels (B,S,80), aux (B,S,128) numpy -> samples (B,steps) float32, steps done, seconds."""
        mels, aux = torch.from_numpy(mels), torch.from_numpy(aux)
        B, S, _ = mels.shape
        S = min(S, max_steps) if max_steps else S
        h1 = torch.zeros(B, RNN_DIMS); h2 = torch.zeros(B, RNN_DIMS); x = torch.zeros(B, 1)
        d = AUX_DIMS
        if self.mode == "RAW":
            U = torch.from_numpy(philox.raw_uniforms(seed, S, B))
        else:
            um, ul = philox.mol_uniforms(seed, S, B)
            UM, UL = torch.from_numpy(um), torch.from_numpy(ul)
        out = []
        t0 = time.perf_counter()
        for i in range(S):
            m_t = mels[:, i, :]
            a1, a2, a3, a4 = (aux[:, i, d * k:d * (k + 1)] for k in range(4))
            x = torch.cat([x, m_t, a1[:, :-1]], dim=1)
            x = self.I(x)
            h1 = self.rnn1(x, h1)
            x = x + h1
            h2 = self.rnn2(torch.cat([x, a2], dim=1), h2)
            x = x + h2
            x = F.relu(self.fc1(torch.cat([x, a3], dim=1)))
            x = F.relu(self.fc2(torch.cat([x, a4], dim=1)))
            logits = self.fc3(x)
            if self.mode == "MOL":
                temp = logits[:, :10] - torch.log(-torch.log(UM[i]))
                k = temp.argmax(dim=1)
                rows = torch.arange(B)
                means = logits[rows, 10 + k]
                ls = torch.clamp(logits[rows, 20 + k], min=float(orc.LOG_SCALE_MIN))
                s = torch.clamp(means + torch.exp(ls) * (torch.log(UL[i]) - torch.log(1. - UL[i])), -1., 1.)
            else:
                p = F.softmax(logits, dim=1)
                k = (torch.cumsum(p, dim=1) < U[i][:, None]).sum(dim=1).clamp(max=self.C - 1)
                s = 2 * k.float() / (self.C - 1.) - 1.
            out.append(s)
            x = s.unsqueeze(-1)
            if time_budget_s and (i & 15) == 15 and time.perf_counter() - t0 > time_budget_s:
                break
        dt = time.perf_counter() - t0
        return torch.stack(out, dim=1).numpy(), len(out), dt


def time_generate(sd, mode, mel_norm, batched, target, overlap, threads=None, time_budget_s=15.0, seed=0):
    """Times a bounded sample of generate(): full conditioning + fold, then as many loop steps as fit
    in the budget.  Returns dict(steps_done, steps_total, folds, seconds_loop, seconds_cond,
    est_total_seconds, out_samples)."""
    port = TorchPort(sd, mode, threads)
    t0 = time.perf_counter()
    mels, aux = orc.upsample_network(mel_norm, sd)
    if batched:
        mels = orc.fold_with_overlap(mels, target, overlap)
        aux = orc.fold_with_overlap(aux, target, overlap)
    else:
        mels, aux = mels[None], aux[None]
    t_cond = time.perf_counter() - t0
    B, S, _ = mels.shape
    port.loop(mels[:, :8], aux[:, :8], seed=seed)               # warm-up
    _, done, dt = port.loop(mels, aux, seed=seed, time_budget_s=time_budget_s)
    est = t_cond + dt * S / done
    return dict(steps_done=done, steps_total=S, folds=B, seconds_loop=dt, seconds_cond=t_cond,
                est_total_seconds=est, out_samples=(mel_norm.shape[1] - 1) * 200, threads=torch.get_num_threads())
