# flake8: noqa
from .initializer_base import InitializerBase
from .provided import BallInitializer, RandomBallInitializer
from .seed import center_of_mass_seeder
