"""RL training experiment (reference: trainers/train.py:27-353).

What is kept: the hyper-parameter file and model directory a later ``ttl_track.py`` run loads
(``<path>/model/hyperparameters.json`` with the reference's keys, ``last_model_state_{actor,critic}.pth``),
the epoch structure of ``rl_train`` (a validation run with the untrained agent, then ``max_ep`` training
episodes of ``Tracker.track_and_train`` with a validation run + model save every ``log_interval``), and
the reference's option names.  What is different and why: subjects come from NIfTI files (``in_odf``,
``in_seed``, ``in_mask``, like ``ttl_track.py``) because the reference's HDF5 datasets need h5py / dwi_ml
(absent); Comet and the Tractometer / oracle validators are not part of the hot
path.  One process per GPU under torchrun: every rank trains on its own seeds with
the learner's gradient all-reduce keeping the replicas in lock-step (algorithms/sac_train.py); rank 0
writes the files.
"""
import json
import os
import random
from os.path import join as pjoin

import numpy as np
import torch

from tracktolearn_b200.environments import NoisyTrackingEnvironment, TrackingEnvironment
from tracktolearn_b200.tracking.tracker import Tracker


class TrackToLearnTraining(object):
    """Reference: trainers/train.py:27 (constructor keys :44-120)."""

    def __init__(self, train_dto):
        self.experiment_path = train_dto['path']
        self.experiment = train_dto['experiment']
        self.name = train_dto['id']
        self.max_ep = train_dto['max_ep']
        self.log_interval = train_dto['log_interval']
        self.noise = train_dto['noise']
        self.lr = train_dto['lr']
        self.gamma = train_dto['gamma']
        self.step_size = train_dto['step_size']
        self.in_odf, self.in_seed, self.in_mask = train_dto['in_odf'], train_dto['in_seed'], train_dto['in_mask']
        self.sh_basis = train_dto.get('sh_basis', 'descoteaux07')
        self.target_sh_order = train_dto.get('target_sh_order', 8)
        self.dataset_file = train_dto.get('dataset_file', self.in_odf)
        self.rng_seed = train_dto['rng_seed']
        self.npv = train_dto['npv']
        self.theta = train_dto['theta']
        self.min_length = train_dto['min_length']
        self.max_length = train_dto['max_length']
        self.binary_stopping_threshold = train_dto['binary_stopping_threshold']
        self.alignment_weighting = train_dto['alignment_weighting']
        self.hidden_dims = train_dto['hidden_dims']
        self.n_actor = train_dto['n_actor']
        self.n_dirs = train_dto['n_dirs']
        self.oracle_checkpoint = train_dto.get('oracle_checkpoint')
        self.oracle_bonus = train_dto.get('oracle_bonus', 0.0) if self.oracle_checkpoint else 0.0
        self.oracle_stopping_criterion = bool(train_dto.get('oracle_stopping_criterion')) and bool(self.oracle_checkpoint)
        self.precision = train_dto.get('precision', 'fp16')
        self.compute_reward = True       # always during training (train.py:96)
        self.last_episode = 0
        if not torch.cuda.is_available():
            raise SystemExit('training (tracktolearn_b200) needs a CUDA device; there is no CPU path')
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        if self.world > 1:
            import torch.distributed as dist
            local = int(os.environ.get('LOCAL_RANK', '0'))
            torch.cuda.set_device(local)
            if not dist.is_initialized():
                dist.init_process_group('nccl', device_id=torch.device('cuda', local))
            self.device = torch.device('cuda', local)
        else:
            self.device = torch.device('cuda', torch.cuda.current_device())
        # every rank draws its own seeds and exploration noise; the weights are broadcast from rank 0
        seed = self.rng_seed + self.rank
        torch.manual_seed(seed)
        np.random.seed(seed)
        self.rng = np.random.RandomState(seed=seed)
        random.seed(seed)
        if self.rank == 0:
            os.makedirs(pjoin(self.experiment_path, 'model'), exist_ok=True)
        # train.py:122-153
        self.hyperparameters = {
            'name': self.name, 'experiment': self.experiment, 'max_ep': self.max_ep,
            'log_interval': self.log_interval, 'lr': self.lr, 'gamma': self.gamma,
            'step_size': self.step_size, 'random_seed': self.rng_seed, 'dataset_file': self.dataset_file,
            'n_seeds_per_voxel': self.npv, 'max_angle': self.theta, 'min_length': self.min_length,
            'max_length': self.max_length, 'binary_stopping_threshold': self.binary_stopping_threshold,
            'experiment_path': self.experiment_path, 'hidden_dims': self.hidden_dims,
            'last_episode': self.last_episode, 'n_actor': self.n_actor, 'n_dirs': self.n_dirs,
            'noise': self.noise, 'alignment_weighting': self.alignment_weighting,
            'oracle_bonus': self.oracle_bonus, 'oracle_checkpoint': self.oracle_checkpoint,
            'oracle_stopping_criterion': self.oracle_stopping_criterion,
        }

    # ------------------------------------------------------------------ files (train.py:155-186)
    def save_hyperparameters(self):
        self.hyperparameters.update({'input_size': self.input_size, 'action_size': self.action_size,
                                     'voxel_size': str(self.voxel_size), 'target_sh_order': self.target_sh_order})
        if self.rank != 0:
            return
        with open(pjoin(self.experiment_path, 'model', 'hyperparameters.json'), 'w') as json_file:
            json_file.write(json.dumps(self.hyperparameters, indent=4, separators=(',', ': ')))

    def save_model(self, alg):
        if self.rank != 0:
            return
        directory = pjoin(self.experiment_path, 'model')
        os.makedirs(directory, exist_ok=True)
        alg.agent.save(directory, 'last_model_state')

    # ------------------------------------------------------------------ environments
    def _env_dto(self, noise):
        return {
            'fa_map': None, 'n_dirs': self.n_dirs, 'step_size': self.step_size, 'theta': self.theta,
            'min_length': self.min_length, 'max_length': self.max_length, 'noise': noise, 'npv': self.npv,
            'rng': self.rng, 'alignment_weighting': self.alignment_weighting, 'oracle_bonus': self.oracle_bonus,
            'oracle_validator': False, 'oracle_stopping_criterion': self.oracle_stopping_criterion,
            'oracle_checkpoint': self.oracle_checkpoint, 'scoring_data': None, 'tractometer_validator': False,
            'binary_stopping_threshold': self.binary_stopping_threshold, 'compute_reward': True,
            'device': self.device, 'target_sh_order': int(self.target_sh_order),
            'in_odf': self.in_odf, 'in_seed': self.in_seed, 'in_mask': self.in_mask, 'sh_basis': self.sh_basis,
            'input_wm': False, 'reference': self.in_mask,
        }

    def get_env(self):
        """Reference: experiment/experiment.py:131-160 (training env: no noise)."""
        return TrackingEnvironment.from_files(self._env_dto(0.0))

    def get_valid_env(self):
        """Reference: experiment/experiment.py:162-187 (validation env, NoisyTrackingEnvironment)."""
        return NoisyTrackingEnvironment.from_files(self._env_dto(self.noise))

    @staticmethod
    def stopping_stats(tractogram):
        """Reference: experiment/experiment.py:292-314: share of every stopping flag."""
        if tractogram is None or len(tractogram) == 0:
            return {}
        from tracktolearn_b200.environments.stopping_criteria import StoppingFlags, is_flag_set
        flags = np.asarray(tractogram.data_per_streamline['flags'])
        return {f.name: float(np.mean(is_flag_set(flags, f))) for f in StoppingFlags}

    # ------------------------------------------------------------------ the loop (train.py:188-330)
    def rl_train(self, alg, env, valid_env):
        i_episode = 0
        t = 0
        train_tracker = Tracker(alg, self.n_actor, prob=0.0, compress=0.0)
        valid_tracker = Tracker(alg, self.n_actor, prob=1.0, compress=0.0, streaming=False)
        log = []

        def validate(episode):
            tractogram, reward = valid_tracker.track_and_validate(valid_env)
            stats = self.stopping_stats(tractogram)
            n = len(tractogram) if tractogram is not None else 0
            if self.rank == 0:
                print('Validation at episode {}: {} streamlines, reward {:.3f}, {}'.format(episode, n, reward, stats))
            self.save_model(alg)
            return {'episode': episode, 'valid_reward': float(reward), 'valid_streamlines': n, 'stopping': stats}
        log.append(validate(i_episode))           # what an untrained network does
        while i_episode < self.max_ep:
            self.last_episode = i_episode
            tractogram, losses, reward, reward_factors = train_tracker.track_and_train(env)
            lengths = tractogram.lengths if tractogram is not None else np.zeros(0)
            avg_length = float(np.mean(lengths)) if len(lengths) else 0.0
            t += int(np.sum(lengths))
            avg_reward = reward / self.n_actor
            if self.rank == 0:
                print('Episode Num: {} Avg len: {:.3f} Avg. reward: {:.3f} sub: {}'.format(
                    i_episode + 1, avg_length, avg_reward, getattr(env, 'subject_id', '')))
            i_episode += 1
            log.append({'episode': i_episode, 'avg_length': avg_length, 'avg_reward': float(avg_reward),
                        'transitions': t, 'losses': {k: float(np.mean(v)) for k, v in losses.items()}})
            if i_episode % self.log_interval == 0:
                log.append(validate(i_episode))
        log.append(validate(i_episode))
        self.training_log = log
        return log

    def run(self):
        """Reference: trainers/train.py:332-353."""
        env = self.get_env()
        valid_env = self.get_valid_env()
        self.input_size = env.get_state_size()
        self.action_size = env.get_action_size()
        self.voxel_size = env.get_voxel_size()
        self.target_sh_order = env.target_sh_order
        alg = self.get_alg(env.max_nb_steps)
        self.save_hyperparameters()
        return self.rl_train(alg, env, valid_env)


def add_training_args(parser):
    """The reference's option names (experiment/experiment.py:383-476, trainers/train.py:356-376); the
    dataset positional is three NIfTI files instead of one HDF5 file."""
    parser.add_argument('path', type=str, help='Path to experiment')
    parser.add_argument('experiment', help='Name of experiment.')
    parser.add_argument('id', type=str, help='ID of experiment.')
    parser.add_argument('in_odf', help='fODF spherical harmonics (.nii / .nii.gz)')
    parser.add_argument('in_seed', help='Seeding mask (.nii / .nii.gz)')
    parser.add_argument('in_mask', help='Tracking mask (.nii / .nii.gz)')
    parser.add_argument('--sh_basis', default='descoteaux07', choices=['descoteaux07', 'tournier07'])
    parser.add_argument('--rng_seed', default=1337, type=int, help='Seed to fix general randomness')
    parser.add_argument('--n_dirs', default=100, type=int, help='Last n steps taken')
    parser.add_argument('--binary_stopping_threshold', type=float, default=0.1,
                        help='Lower limit for interpolation of tracking mask value.')
    parser.add_argument('--n_actor', default=4096, type=int, help='Number of learners')
    parser.add_argument('--hidden_dims', default='1024-1024-1024', type=str, help='Hidden layers of the model')
    parser.add_argument('--max_ep', default=1000, type=int, help='Number of episodes to run the training algorithm')
    parser.add_argument('--log_interval', default=50, type=int, help='Validate and save the model every n episodes')
    parser.add_argument('--lr', default=0.0005, type=float, help='Learning rate')
    parser.add_argument('--gamma', default=0.95, type=float, help='Gamma param for reward discounting')
    parser.add_argument('--alignment_weighting', default=1, type=float, help='Alignment weighting for reward')
    parser.add_argument('--npv', default=2, type=int, help='Number of random seeds per seeding mask voxel.')
    parser.add_argument('--theta', default=30, type=int, help='Max angle between segments for tracking.')
    parser.add_argument('--min_length', type=float, default=20., metavar='m', help='Minimum length of a streamline in mm.')
    parser.add_argument('--max_length', type=float, default=200., metavar='M', help='Maximum length of a streamline in mm.')
    parser.add_argument('--step_size', default=0.75, type=float, help='Step size for tracking')
    parser.add_argument('--noise', default=0.0, type=float, metavar='sigma', help='Noise added to the actions at validation')
    parser.add_argument('--oracle_checkpoint', type=str, default=None, help='Checkpoint file (.ckpt) of the Oracle')
    parser.add_argument('--oracle_stopping_criterion', action='store_true', help='Stop streamlines the oracle rejects')
    parser.add_argument('--oracle_bonus', default=10, type=float, help='Sparse oracle bonus')
    parser.add_argument('--precision', default='fp16', choices=['fp16', 'tf32', 'bf16', 'fp32'],
                        help='Arithmetic of the rollout actor (the learner is fp32 like the reference)')
