import sys, time, json, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from tracktolearn_b200 import _lib, synthetic
from tracktolearn_b200.algorithms.sac_auto import SACAuto
from tracktolearn_b200.datasets.utils import MRIDataVolume
from tracktolearn_b200.environments import NoisyTrackingEnvironment
dev = torch.device('cuda:0')
sub = synthetic.make_subject(bench.SHAPE, seed=1234, device=dev, with_peaks=False)
affine = np.diag([bench.VOXEL_MM]*3+[1.0])
subject = (MRIDataVolume(sub['sh'], affine), MRIDataVolume(sub['mask'], affine), MRIDataVolume(sub['seed_mask'], affine), None, affine)
dto = {'n_dirs': 100, 'theta': 30.0, 'npv': 1, 'binary_stopping_threshold': 0.1, 'step_size': bench.STEP_MM, 'min_length': 10.0, 'max_length': 300.0,
       'oracle_checkpoint': None, 'oracle_stopping_criterion': False, 'scoring_data': None, 'compute_reward': False, 'alignment_weighting': 0.0, 'oracle_bonus': 0.0,
       'rng': np.random.RandomState(1337), 'device': dev, 'target_sh_order': 8, 'noise': 0.0, 'fa_map': None, 'state_of_stopped': False}
env = NoisyTrackingEnvironment(subject, 'testing', dto)
env.seeds = bench.draw_seeds(sub['seed_mask'].cpu().numpy(), 0)[:800000]
alg = SACAuto(615, 3, bench.HIDDEN, n_actors=50000, device=dev, precision='bf16')
alg.agent.actor.load_state_dict(synthetic.actor_state_dict(615, bench.HIDDEN, seed=1111, kind='tracking'))
def sync(): torch.cuda.synchronize()
for rep in range(6):
    alg.use_cuda_graph = rep >= 3
    sync(); t0 = time.perf_counter()
    st = env.reset_streaming(0, len(env.seeds), 50000, fp32_state=False); sync(); t1 = time.perf_counter()
    counts = []
    def on_step(it):
        pass
    alg.sync_every = 8
    alg.validation_episode(st, env, 0.0); sync(); t2 = time.perf_counter()
    tr = env.get_streamlines(copy=False); sync(); t3 = time.perf_counter()
    steps = env.streamline_steps()
    print('graph' if alg.use_cuda_graph else 'plain', 'rep', rep, 'reset %.1f ms, episode %.1f ms (%d env steps), get_streamlines %.1f ms, total %.1f ms, units %d -> %.1f M/s' % (
        1e3*(t1-t0), 1e3*(t2-t1), alg.last_episode_steps, 1e3*(t3-t2), 1e3*(t3-t0), steps, steps/(t3-t0)/1e6))
# where does the episode time go: full-occupancy part vs tail
st = env.reset_streaming(0, len(env.seeds), 50000, fp32_state=False); sync()
actor = alg.agent.actor
buf = torch.empty((50000,3), device=dev)
t0=time.perf_counter(); it=0; log=[]
while True:
    actor.forward_device(None, 0.0, n_rows_dev=env.alive_count_tensor(), n_rows=env._n_alive_host, want_logp=False, out_action=buf, state_bf16=env.current_state_bf16(), layout=env.bf16_layout)
    env.step_device(buf); env.harvest_device(); it+=1
    if it % 8 == 0:
        n = env.n_alive(); log.append((it, n, time.perf_counter()-t0))
        if env._n_alive_host == 0: break
full=[l for l in log if l[1]==50000]
print('steps at full occupancy: %d in %.1f ms; tail: %d steps in %.1f ms' % (full[-1][0], 1e3*full[-1][2], log[-1][0]-full[-1][0], 1e3*(log[-1][2]-full[-1][2])))
print([ (a,b) for a,b,_ in log[len(full)::4]])
