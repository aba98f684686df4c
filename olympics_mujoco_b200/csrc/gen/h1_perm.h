#pragma once
// GENERATED: qpos/qvel address of UnitreeH1 observation-spec entry k (UnitreeH1.py:303-355 minus the arm joints)
__device__ constexpr int OM_H1_PERM[17] = {0, 1, 2, 3, 4, 5, 16, 13, 12, 11, 14, 15, 8, 7, 6, 9, 10};
static const int OM_H1_PERM_HOST[17] = {0, 1, 2, 3, 4, 5, 16, 13, 12, 11, 14, 15, 8, 7, 6, 9, 10};
