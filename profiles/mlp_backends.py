import sys, os, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import merging_gym_b200 as mg
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'dqn_policies.npz'))
sd = {k.split('/',1)[1]: z[k] for k in z.files if k.startswith('L1_1445/') and 'traj' not in k and 'result' not in k}
for n in (128, 1000, 5000, 1<<18):
    env = mg.MergeVecEnv(n, seed=3); env.rollout(120); obs = env.step(*env.sample_actions())[0].clone()
    for name, sdict in (("ckpt", sd), ("rand", None)):
        f = mg.MLPPolicy(10, 5, state_dict=sdict, seed=7)
        tc = mg.MLPPolicy(10, 5, state_dict=f.state_dict(), backend=os.environ.get("TC", "tf32x3"))
        qf = torch.empty(n, 5, device='cuda'); qt = torch.empty(n, 5, device='cuda')
        af = f.act(obs, q_out=qf); at = tc.act(obs, q_out=qt)
        torch.cuda.synchronize()
        w = {k: v.double().cpu() for k, v in f.state_dict().items()}
        x = obs.double().cpu()
        h = torch.relu(x @ w['fc1.weight'].t() + w['fc1.bias']); h = torch.relu(h @ w['fc2.weight'].t() + w['fc2.bias'])
        q64 = h @ w['out.weight'].t() + w['out.bias']
        sc = q64.abs().max().item()
        print(n, name, "scale %.3g" % sc, "err fused %.3e  err tc %.3e" % ((qf.double().cpu()-q64).abs().max().item()/sc, (qt.double().cpu()-q64).abs().max().item()/sc),
              "action agreement tc/fused %.6f  tc/fp64 %.6f fused/fp64 %.6f" % ((af==at).float().mean().item(), (at.cpu().long()==q64.argmax(1)).float().mean().item(), (af.cpu().long()==q64.argmax(1)).float().mean().item()))
n = 1 << 18
act = torch.empty(n, dtype=torch.uint8, device='cuda')
for be in ("fused", "tf32x3"):
    p = mg.MLPPolicy(10, 5, state_dict=sd, backend=be)
    for _ in range(5): p.act(obs, out=act)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): p.act(obs, out=act)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(be, "ms", ms, "TFLOP/s (fp32-equivalent)", n * 45000 / (ms * 1e-3) / 1e12)
