"""TicTacToe plug-in (config values and rules of the reference's ``games/tictactoe.py``)."""
import numpy

from ._boards import BoardGame, BoardVector
from ._config import BaseMuZeroConfig
from .abstract_game import AbstractGame


class MuZeroConfig(BaseMuZeroConfig):
    _NAME = "tictactoe"
    _OVERRIDES = dict(
        observation_shape=(3, 3, 3), action_space=list(range(9)), players=list(range(2)),
        opponent="expert", max_moves=9, num_simulations=25, discount=1,
        root_dirichlet_alpha=0.1,
        network="resnet", blocks=1, channels=16,
        reduced_channels_reward=16, reduced_channels_value=16, reduced_channels_policy=16,
        resnet_fc_reward_layers=[8], resnet_fc_value_layers=[8], resnet_fc_policy_layers=[8],
        encoding_size=32, fc_value_layers=[], fc_policy_layers=[],
        training_steps=1000000, batch_size=64, value_loss_weight=0.25, lr_init=0.003,
        lr_decay_rate=1, lr_decay_steps=10000, replay_buffer_size=3000, num_unroll_steps=20,
        td_steps=20, ratio=None,
    )
    _TEMPERATURE_SCHEDULE = ((None, 1),)


class TicTacToeVector(BoardVector):
    H = W = K = 3
    OBS_DTYPE = numpy.int32
    REWARD_SCALE = 20          # games/tictactoe.py:144


class Game(BoardGame, AbstractGame):
    VECTOR = TicTacToeVector

    def action_to_string(self, action_number):
        return f"Play row {action_number // 3 + 1}, column {action_number % 3 + 1}"
