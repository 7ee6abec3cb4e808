from .MMETrainingModule import MMETrainingModule
from .SimpleTrain import SimpleTrainModule
from .TrainingBase import TrainingBase, getClassWeight

__all__ = ["TrainingBase", "SimpleTrainModule", "MMETrainingModule", "getClassWeight"]
