"""SAC with automatic entropy adjustment: the training runner
(reference: trainers/sac_auto_train.py:19-129).

    python -m tracktolearn_b200.trainers.sac_auto_train PATH EXPERIMENT ID in_odf in_seed in_mask [options]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        -m tracktolearn_b200.trainers.sac_auto_train ...          # N GPUs, gradient all-reduce
"""
import argparse
from argparse import RawTextHelpFormatter

from tracktolearn_b200.algorithms.sac_auto import SACAuto
from tracktolearn_b200.trainers.train import TrackToLearnTraining, add_training_args


class SACAutoTrackToLearnTraining(TrackToLearnTraining):
    """Reference: trainers/sac_auto_train.py:19-76."""

    def __init__(self, sac_auto_train_dto):
        super().__init__(sac_auto_train_dto)
        self.alpha = sac_auto_train_dto['alpha']
        self.batch_size = sac_auto_train_dto['batch_size']
        self.replay_size = sac_auto_train_dto['replay_size']
        self.start_timesteps = sac_auto_train_dto.get('start_timesteps')

    def save_hyperparameters(self):
        self.hyperparameters.update({'algorithm': 'SACAuto', 'alpha': self.alpha, 'batch_size': self.batch_size,
                                     'replay_size': self.replay_size})
        super().save_hyperparameters()

    def get_alg(self, max_nb_steps):
        alg = SACAuto(self.input_size, self.action_size, self.hidden_dims, self.lr, self.gamma, self.alpha,
                      self.n_actor, self.batch_size, self.replay_size, self.rng, self.device,
                      precision=self.precision)
        alg.enable_training(lr=self.lr, gamma=self.gamma, replay_size=self.replay_size, batch_size=self.batch_size,
                            start_timesteps=self.start_timesteps)
        return alg


def add_sac_auto_args(parser):
    parser.add_argument('--alpha', default=0.2, type=float, help='Initial temperature parameter')
    parser.add_argument('--batch_size', default=2 ** 12, type=int,
                        help='How many tuples to sample from the replay buffer.')
    parser.add_argument('--replay_size', default=1e6, type=int, help='How many tuples to store in the replay buffer.')
    parser.add_argument('--start_timesteps', default=None, type=int,
                        help='Transitions to gather before the first update [80000, algorithms/sac_auto.py:131]')


def parse_args(argv=None):
    """ Train a tracking agent with SAC (automatic entropy adjustment). """
    parser = argparse.ArgumentParser(description=parse_args.__doc__, formatter_class=RawTextHelpFormatter)
    add_training_args(parser)
    add_sac_auto_args(parser)
    return parser.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    experiment = SACAutoTrackToLearnTraining(vars(args))
    experiment.run()
    return experiment


if __name__ == '__main__':
    main()
