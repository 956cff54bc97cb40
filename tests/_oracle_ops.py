"""Re-export of the oracle-backed CPU autograd operators (oracle/torch_ops.py) for the tests."""
from oracle.torch_ops import *  # noqa: F401,F403
from oracle.torch_ops import (OracleDCT, OracleSelectiveScan, OracleSelectiveScanF32, make_pair,  # noqa: F401
                              oracleize)
