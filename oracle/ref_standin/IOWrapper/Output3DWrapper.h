// IOWrapper/Output3DWrapper.h — STUB of a reference header (see ../Eigen/Core): FullSystem/CoarseTracker.h only passes
// pointers to it; the real header needs Eigen::aligned_allocator maps and the viewer types.
#pragma once
namespace dso { namespace IOWrap { class Output3DWrapper; } }
