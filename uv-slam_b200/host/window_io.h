// window_io.h — writer of the `uvs_window v1` file format (uv-slam_b200/window.py) from the C struct the library takes.
//
// the reference ships no recorded windows, so real-data parity fixtures have to be dumped from a
// running vins_estimator.  A patched Estimator::optimization() assembles a UvsWindow from the same
// members it hands to Ceres (estimator.cpp:776-934); GpuWindowProblem::save() / uvs_host_save_window() write exactly
// the bytes Window.to_bytes() writes, so a dump loads in tests/ and bench.py like a synthetic window.
#pragma once
#include "../../include/uvs.h"

namespace uvs_host {
// returns UVS_OK, UVS_ERR_INVALID_ARG (null window / path, unknown prior block kind) or UVS_ERR_UNSUPPORTED (file cannot be written)
int save_window(const UvsWindow &w, const char *path);
}  // namespace uvs_host

extern "C" int uvs_host_save_window(const UvsWindow *w, const char *path);
