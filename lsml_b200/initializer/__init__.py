# flake8: noqa
from .initializer_base import InitializerBase
from .provided import BallInitializer, RandomBallInitializer
