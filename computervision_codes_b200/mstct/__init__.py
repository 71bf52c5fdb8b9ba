from .encoder import (GLRBlock, Global_Relational_Block, Local_Relational_Block, TemporalEncoder,  # noqa: F401
                      Temporal_Merging_Block)
from .mixer import Temporal_Mixer, linear_layer  # noqa: F401
from .network import Classifier, VideoNas  # noqa: F401
