"""uv-slam_b200 — B200-native sliding-window backend solve for UV-SLAM.

Python is only the test / bench harness here: the product is `libuvs_b200.so` (hand-written
sm_100a CUDA behind the C ABI of include/uvs.h) plus the C++ drop-in classes in host/.
"""
from .window import (Window, UvsWindowStruct, UvsOptionsStruct, UvsSummaryStruct, UvsPriorStruct,  # noqa: F401
                     default_options, window_array)
from .binding import Solver, load_library, library_path, UvsError  # noqa: F401
