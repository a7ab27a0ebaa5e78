"""W1: ActorCritic.save / load (reference algorithms/shared/offpolicy.py:327-357): both files, the
reference's key layout, round trip before and after training."""
import os

import numpy as np
import pytest
import torch

from tracktolearn_b200 import synthetic

pytestmark = pytest.mark.gpu

HIDDEN = '128-128-128'
ACTOR_KEYS = {'layers.0.weight': (128, 615), 'layers.0.bias': (128,), 'layers.2.weight': (128, 128),
              'layers.2.bias': (128,), 'layers.4.weight': (128, 128), 'layers.4.bias': (128,),
              'layers.6.weight': (6, 128), 'layers.6.bias': (6,)}
CRITIC_KEYS = {}
for q in ('q1', 'q2'):
    CRITIC_KEYS.update({q + '.0.weight': (128, 618), q + '.0.bias': (128,), q + '.2.weight': (128, 128),
                        q + '.2.bias': (128,), q + '.4.weight': (128, 128), q + '.4.bias': (128,),
                        q + '.6.weight': (1, 128), q + '.6.bias': (1,)})


def _agent(precision='fp16'):
    from tracktolearn_b200.algorithms.sac_auto import SACAuto
    return SACAuto(615, 3, HIDDEN, n_actors=64, device=torch.device('cuda:0'), precision=precision)


def _check_files(path):
    a = torch.load(os.path.join(path, 'last_model_state_actor.pth'), map_location='cpu')
    c = torch.load(os.path.join(path, 'last_model_state_critic.pth'), map_location='cpu')
    assert {k: tuple(v.shape) for k, v in a.items()} == ACTOR_KEYS
    assert {k: tuple(v.shape) for k, v in c.items()} == CRITIC_KEYS
    assert all(v.dtype == torch.float32 and v.device.type == 'cpu' for v in list(a.values()) + list(c.values()))
    return a, c


def test_save_load_round_trip_untrained_and_loaded(tmp_path):
    alg = _agent()
    alg.agent.actor.load_state_dict(synthetic.actor_state_dict(615, HIDDEN, seed=5, kind='tracking'))
    alg.agent.save(str(tmp_path), 'last_model_state')            # never trained: fresh critic, like the reference
    a, c = _check_files(str(tmp_path))
    other = _agent()
    other.agent.load(str(tmp_path), 'last_model_state')
    st = torch.randn((200, 615), device='cuda')
    assert torch.equal(alg.agent.select_action(st, 0.0), other.agent.select_action(st, 0.0))
    for k, v in other.agent.state_dict()[1].items():
        assert torch.equal(v.cpu(), c[k])
    # what was loaded is what is saved again
    d2 = tmp_path / 'again'
    d2.mkdir()
    other.agent.save(str(d2), 'last_model_state')
    a2, c2 = _check_files(str(d2))
    assert all(torch.equal(a[k], a2[k]) for k in a) and all(torch.equal(c[k], c2[k]) for k in c)
    # a missing critic file raises, as in the reference
    os.remove(os.path.join(str(d2), 'last_model_state_critic.pth'))
    with pytest.raises(FileNotFoundError):
        _agent().agent.load(str(d2), 'last_model_state')


def test_trained_agent_saves_the_learners_critic_and_resumes(tmp_path):
    alg = _agent()
    # a checkpoint's critic must survive enable_training
    ck = synthetic.critic_state_dict(615, HIDDEN, seed=77)
    alg.agent.load_state_dict((synthetic.actor_state_dict(615, HIDDEN, seed=5, kind='tracking'), ck))
    learner = alg.enable_training(replay_size=4096, batch_size=64, start_timesteps=1)
    for k, v in learner.critic.state_dict().items():
        assert torch.equal(v.cpu(), ck[k])
    g = torch.Generator(device='cuda').manual_seed(0)
    batch = (torch.randn((64, 615), device='cuda', generator=g), torch.rand((64, 3), device='cuda', generator=g) * 2 - 1,
             torch.randn((64, 615), device='cuda', generator=g), torch.rand((64,), device='cuda', generator=g),
             torch.ones((64,), device='cuda'))
    for _ in range(3):
        alg.update(batch)
    alg.agent.save(str(tmp_path), 'last_model_state')
    a, c = _check_files(str(tmp_path))
    for k, v in learner.critic.state_dict().items():
        assert torch.equal(v.cpu(), c[k])
    assert any(not torch.equal(c[k], ck[k]) for k in c)              # the updates moved the critic
    for k, v in learner.actor.state_dict().items():
        assert torch.equal(v.cpu(), a[k])
    # resume: a new agent loads both, starts training, and its learner holds the saved networks
    resumed = _agent()
    resumed.agent.load(str(tmp_path), 'last_model_state')
    l2 = resumed.enable_training(replay_size=4096, batch_size=64, start_timesteps=1)
    for k, v in l2.critic.state_dict().items():
        assert torch.equal(v.cpu(), c[k])
    for k, v in l2.target_critic.state_dict().items():
        assert torch.equal(v.cpu(), c[k])
    st = torch.randn((100, 615), device='cuda')
    assert torch.equal(alg.agent.select_action(st, 0.0), resumed.agent.select_action(st, 0.0))
    # loading into an agent that is already training goes through the learner
    resumed.agent.load(str(tmp_path), 'last_model_state')
    assert torch.equal(alg.agent.select_action(st, 0.0), resumed.agent.select_action(st, 0.0))
