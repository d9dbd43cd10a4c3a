# flake8: noqa
from .ball import BallInitializer, RandomBallInitializer, ThresholdBallInitializer
from .threshold import ThresholdInitializer
