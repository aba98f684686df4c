"""Compile the two in-scope MJCF files into the JSON tables shipped with the package.

Run in the authoring container (the only place /root/reference exists):
    python tools/gen_models.py [/root/reference]
Outputs olympics_mujoco_b200/models/{unitree_h1,unitree_h1_arms,stick_figure_a3}.json.
The JSON holds numeric tables only (no reference source text); the GPU box loads these.
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from olympics_mujoco_b200 import mjcf  # noqa: E402

ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
data = ref / "olympic_mujoco" / "environments" / "data"
out = ROOT / "olympics_mujoco_b200" / "models"
out.mkdir(exist_ok=True)

h1 = mjcf.compile_unitree_h1(data / "unitree_h1" / "h1.xml")
h1.save_json(out / "unitree_h1.json")
h1_arms = mjcf.compile_unitree_h1(data / "unitree_h1" / "h1.xml", disable_arms=False)
h1_arms.name = "UnitreeH1_arms"
h1_arms.save_json(out / "unitree_h1_arms.json")
a3 = mjcf.compile_stick_figure_a3(data / "stickFigure_A3" / "a3.xml")
a3.save_json(out / "stick_figure_a3.json")
for m in (h1, h1_arms, a3):
    print(f"{m.name}: nbody={m.nbody} njnt={m.njnt} nq={m.nq} nv={m.nv} nsite={m.nsite} "
          f"nu={m.nu} mass={m.total_mass:.4f}")
