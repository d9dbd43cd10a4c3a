# flake8: noqa
from .initializer_base import InitializerBase
from .provided import (BallInitializer, RandomBallInitializer, ThresholdBallInitializer,
                       ThresholdInitializer)
from .seed import center_of_mass_seeder
