#pragma once
// GENERATED: body / site ids the WalkingTask reads (StickFigureA3.py:95-103, walking_task.py:254-263) and the model's total mass (mj_getTotalmass)
constexpr int OM_A3_ROOT = 1;
constexpr int OM_A3_HEAD = 2;
constexpr int OM_A3_LFOOT = 10;
constexpr int OM_A3_RFOOT = 7;
constexpr int OM_A3_LSITE = 1;
constexpr int OM_A3_RSITE = 0;
constexpr float OM_A3_TOTAL_MASS = 4.082136e+01f;
