from .network import (VideoNas, FPN, BaseCausalTCN, Refinement, DilatedResidualCausalLayer,  # noqa: F401
                      DilatedResidualLayer)
