import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
import merging_gym_b200 as mg
n=1<<20
env=mg.MergeVecEnv(n, episode_info=False); env.rollout(300)
a1,a2=env.sample_actions(); h1=a1.cpu().numpy().copy(); h2=a2.cpu().numpy().copy()
for zc in (False, True):
    for _ in range(3): env.step_host(h1,h2,zero_copy=zc)
    t=time.perf_counter()
    for _ in range(30): env.step_host(h1,h2,zero_copy=zc)
    dt=(time.perf_counter()-t)/30
    print("zero_copy",zc,"ms/step",dt*1e3,"steps/s",n/dt, "GB/s", n*52/dt/1e9)
