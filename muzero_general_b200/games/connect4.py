"""Connect4 plug-in (config values and rules of the reference's ``games/connect4.py``)."""
import numpy

from ._boards import BoardGame, BoardVector
from ._config import BaseMuZeroConfig
from .abstract_game import AbstractGame


class MuZeroConfig(BaseMuZeroConfig):
    _NAME = "connect4"
    _OVERRIDES = dict(
        observation_shape=(3, 6, 7), action_space=list(range(7)), players=list(range(2)),
        opponent="expert", max_moves=42, num_simulations=200, discount=1,
        root_dirichlet_alpha=0.3,
        network="resnet", blocks=3, channels=64,
        reduced_channels_reward=2, reduced_channels_value=2, reduced_channels_policy=4,
        resnet_fc_reward_layers=[64], resnet_fc_value_layers=[64], resnet_fc_policy_layers=[64],
        encoding_size=32, fc_dynamics_layers=[64], fc_reward_layers=[64],
        fc_value_layers=[], fc_policy_layers=[],
        training_steps=100000, batch_size=64, value_loss_weight=0.25, lr_init=0.005,
        lr_decay_rate=1, lr_decay_steps=10000, replay_buffer_size=10000, num_unroll_steps=42,
        td_steps=42, ratio=None,
    )
    _TEMPERATURE_SCHEDULE = ((None, 1),)


class Connect4Vector(BoardVector):
    H, W, K = 6, 7, 4
    GRAVITY = True
    OBS_DTYPE = numpy.float64
    REWARD_SCALE = 10          # games/connect4.py:144


class Game(BoardGame, AbstractGame):
    VECTOR = Connect4Vector

    def action_to_string(self, action_number):
        return f"Play column {action_number + 1}"
