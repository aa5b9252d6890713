#!/usr/bin/env python
"""f16x3 accuracy over weight and input scales: Q error against an fp64 evaluation, relative to the largest |Q|, for
networks whose weights are scaled by 1e-3 ... 10 and inputs by 1e-2 ... 1e3 (the power-of-two operand scaling and the
hi / lo split have to hold up everywhere short of fp16's range: inputs below 65504, hidden-1 activations below 5.2e5)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import merging_gym_b200 as mg  # noqa: E402

n = 20000
g = torch.Generator().manual_seed(0)
rows = []
for wscale in (1e-3, 1e-2, 1.0, 10.0):
    for xscale in (1e-2, 1.0, 1e2, 1e3):
        sd = {"fc1.weight": (torch.rand(200, 10, generator=g) * 2 - 1) * wscale, "fc1.bias": (torch.rand(200, generator=g) * 2 - 1) * wscale,
              "fc2.weight": (torch.rand(100, 200, generator=g) * 2 - 1) * wscale, "fc2.bias": (torch.rand(100, generator=g) * 2 - 1) * wscale,
              "out.weight": torch.rand(5, 100, generator=g) * 2 - 1, "out.bias": torch.rand(5, generator=g)}
        obs = ((torch.rand(n, 10, generator=g) * 2 - 1) * xscale).cuda().contiguous()
        x = obs.double().cpu()
        h = torch.relu(x @ sd["fc1.weight"].double().t() + sd["fc1.bias"].double())
        h1max = h.max().item()
        h = torch.relu(h @ sd["fc2.weight"].double().t() + sd["fc2.bias"].double())
        ref = h @ sd["out.weight"].double().t() + sd["out.bias"].double()
        scale = ref.abs().max().item()
        row = {"weight_scale": wscale, "input_scale": xscale, "max_hidden1": h1max}
        for be in ("fused", "tf32x3", "f16x3"):
            p = mg.MLPPolicy(10, 5, state_dict=sd, backend=be)
            q = torch.empty(n, 5, device="cuda")
            p.act(obs, q_out=q)
            row[be] = (q.double().cpu() - ref).abs().max().item() / scale
        rows.append(row)
        print(json.dumps(row), flush=True)
worst = max(r["f16x3"] for r in rows)
print(json.dumps({"worst_f16x3": worst, "worst_tf32x3": max(r["tf32x3"] for r in rows), "worst_fp32": max(r["fused"] for r in rows)}))
