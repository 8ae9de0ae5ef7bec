// FullSystem/FullSystem.h — STUB (see ../Eigen/Core): FullSystem/ResidualProjections.h includes it without using anything
// of it in the functions compiled by `make ref` (projectPoint, derive_idepth).
#pragma once
