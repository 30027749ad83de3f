from .AWGNPassedDatagen import AWGNPassedDatagen
from .ConnectingMatrix import ConnectingMatrix
from .ConnectingMatrixTorch import ConnectingMatrixTorch
from .Functions import Functions
