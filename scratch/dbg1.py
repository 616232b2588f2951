import sys, numpy as np, torch
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from dl_reference_models_b200 import maps
from dl_reference_models_b200.batched_env import BatchedMapfEnv
from oracle.oracle import OracleBatch
from gpu_utils import gpu_channels
cfg = {"num_agents": 4, "sensor_range": 2, "steps_per_episode": 100, "seed": 123}
grid=maps.get_grid("ReferenceModel-2-1"); cfg["grid"]=grid
B=4096; seed=999
ob=OracleBatch(cfg,grid,B,seed=seed); env=BatchedMapfEnv(cfg,num_envs=B)
rng=np.random.default_rng(seed)
ob.reset(2); st=ob.state(); out=env.reset(starts=st["starts"],goals=st["goals"])
e=820
print("start",st["positions"][e].tolist(),"goals",st["goals"][e].tolist())
for s in range(5):
    acts=rng.integers(0,5,(B,4)).astype(np.int8)
    ob.step(acts); out=env.step(torch.from_numpy(acts)); got=gpu_channels(env,out)
    st=ob.state()
    print("step",s,"acts",acts[e].tolist())
    print(" ref pos",st["positions"][e].tolist(),"reached",st["reached"][e].tolist(),"blocking",ob.buf["blocking"][e].tolist(),"bp_prev_out",ob.buf["blocking_prev"][e].tolist(),"state bp",st["blocking_prev"][e].tolist(), "intended", ob.buf["intended_next"][e].tolist())
    print(" gpu pos",got["positions"][e].tolist(),"reached",got["reached"][e].tolist(),"blocking",got["blocking"][e].tolist(),"bp_prev_out",got["blocking_prev"][e].tolist(),"aflags",env.state["agent_flags"][e].tolist())
    bad=np.argwhere(got["blocking_prev"]!=ob.buf["blocking_prev"]); print(" bad",bad.tolist()[:5])
