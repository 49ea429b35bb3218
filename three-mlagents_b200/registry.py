"""Task registry — same public surface as the reference's backend/mlagents/registry.py
(`TaskSpec`, `TASKS`, `list_tasks`, `list_task_cards`, `get_task`, `make_env`; registry.py:18-370).

All 19 task cards are kept so `three-mlagents list/inspect` and any caller of `get_task` keep
working.  Only the four tasks on the hot path (basic, ball3d, gridworld, push) and walljump, brickbreak, bicycle, glider have a CUDA env
factory here; the reference's other Gymnasium tasks are listed with `env_factory=None`, so — by the
reference's own rule `trainable = interface == "gymnasium" and env_factory is not None`
(registry.py:41-43) — they report `trainable: false` in this backend and `make_env` raises the same
ValueError the reference raises for its roadmap tasks (registry.py:368-369).
"""
from __future__ import annotations

from dataclasses import asdict, dataclass, field
from typing import Any, Callable, Literal

from . import envs

Interface = Literal["gymnasium", "pettingzoo", "mlagents-llapi", "external"]
ResearchTier = Literal["foundation", "benchmark", "frontier", "roadmap"]


@dataclass(frozen=True)
class TaskSpec:
    id: str
    title: str
    family: str
    interface: Interface
    research_tier: ResearchTier
    default_algorithm: str
    policy_prefix: str
    total_timesteps: int
    eval_episodes: int = 20
    n_envs: int = 1
    reward_threshold: float | None = None
    tags: tuple[str, ...] = ()
    observation: str = "vector"
    action: str = "discrete"
    publication_role: str = "supporting"
    status: str = "standardized"
    notes: str = ""
    env_factory: Callable[[], Any] | None = field(default=None, repr=False, compare=False)

    @property
    def trainable(self) -> bool:
        return self.interface == "gymnasium" and self.env_factory is not None

    def card(self) -> dict[str, Any]:
        data = asdict(self)
        data.pop("env_factory", None)
        data["trainable"] = self.trainable
        return data


_NO_CUDA = "Gymnasium task of the reference; no CUDA backend in three-mlagents_b200 yet"

# id, title, family, interface, tier, algo, timesteps, eval_eps, n_envs, threshold, tags, extras
_TABLE: list[tuple] = [
    ("basic", "Basic Move-To-Goal", "control", "gymnasium", "foundation", "dqn", 25_000, 50, 1, 0.85,
     ("sparse-reward", "tabular-state", "unity-ml-agents"),
     dict(publication_role="unit sanity check for action/observation plumbing", env_factory=envs.make_basic_env)),
    ("ball3d", "3D Ball Balance", "continuous-control", "gymnasium", "foundation", "ppo", 150_000, 30, 8, 150.0,
     ("physics", "stability", "unity-ml-agents"),
     dict(publication_role="browser/Unity parity smoke benchmark", env_factory=envs.make_ball3d_env)),
    ("gridworld", "GridWorld Goal-Conditioned Navigation", "navigation", "gymnasium", "foundation", "dqn", 100_000, 100, 1, 0.75,
     ("goal-conditioned", "procedural-layout", "discrete-control"),
     dict(publication_role="generalization and seed-control baseline", env_factory=envs.make_gridworld_env)),
    ("push", "Push Block", "navigation", "gymnasium", "benchmark", "dqn", 200_000, 100, 1, 0.65,
     ("object-manipulation", "sparse-reward", "planning"),
     dict(publication_role="single-agent manipulation transfer task", env_factory=envs.make_push_env)),
    ("walljump", "Wall Jump", "navigation", "gymnasium", "benchmark", "dqn", 150_000, 100, 1, 0.7,
     ("conditional-skill", "exploration", "procedural-wall"),
     dict(publication_role="conditional-control benchmark", env_factory=envs.make_walljump_env)),
    ("brickbreak", "Brick Break", "arcade", "gymnasium", "benchmark", "ppo", 500_000, 50, 8, None,
     ("arcade", "partial-observability-lite", "long-horizon"),
     dict(publication_role="small arcade control benchmark before ALE/Procgen", env_factory=envs.make_brick_break_env)),
    ("bicycle", "Bicycle Balance and Navigation", "continuous-control", "gymnasium", "benchmark", "ppo", 500_000, 50, 8, None,
     ("underactuated-control", "stability", "navigation"),
     dict(publication_role="control-system benchmark", env_factory=envs.make_bicycle_env)),
    ("glider", "Dynamic Soaring Glider", "aerospace", "gymnasium", "frontier", "ppo", 1_000_000, 50, 8, None,
     ("aerodynamics", "energy-management", "long-horizon"),
     dict(publication_role="domain-specific continuous physics case study", env_factory=envs.make_glider_env)),
    ("labyrinth", "Labyrinth / NetHack-Inspired Navigation", "games", "gymnasium", "frontier", "ppo", 2_000_000, 100, 8, None,
     ("pixels", "maze", "memory", "exploration"),
     dict(observation="image", publication_role="first serious game-like benchmark in this repo", notes=_NO_CUDA)),
    ("astrodynamics", "Orbital Rendezvous and Docking", "aerospace", "gymnasium", "frontier", "ppo", 2_000_000, 50, 8, None,
     ("orbital-mechanics", "safety", "long-horizon"),
     dict(publication_role="physics-heavy scientific case study", notes=_NO_CUDA)),
    ("kraken", "Kraken Fleet Combat", "games", "gymnasium", "benchmark", "ppo", 1_000_000, 50, 8, None,
     ("multi-unit-control", "coordination", "combat"),
     dict(action="multi-discrete", publication_role="compact multi-unit control benchmark", notes=_NO_CUDA)),
    ("ant", "MuJoCo Ant", "continuous-control", "gymnasium", "benchmark", "ppo", 3_000_000, 20, 8, None,
     ("mujoco", "locomotion", "external-standard"),
     dict(action="continuous", publication_role="external control baseline", notes=_NO_CUDA)),
    ("worm", "MuJoCo Swimmer / Worm", "continuous-control", "gymnasium", "benchmark", "ppo", 2_000_000, 20, 8, None,
     ("mujoco", "locomotion", "external-standard"),
     dict(action="continuous", publication_role="external control baseline", notes=_NO_CUDA)),
    ("foodcollector", "Food Collector", "multi-agent", "pettingzoo", "roadmap", "ippo", 2_000_000, 20, 1, None,
     ("multi-agent", "mixed-action", "competitive-cooperative"),
     dict(action="hybrid", publication_role="PettingZoo conversion target",
          status="needs PettingZoo ParallelEnv wrapper before paper-grade training",
          notes="Do not force through single-agent SB3; use PettingZoo plus SuperSuit/RLlib/CleanRL IPPO/MAPPO.")),
    ("intersection", "Traffic Intersection", "multi-agent", "pettingzoo", "frontier", "mappo", 5_000_000, 20, 1, None,
     ("multi-agent", "safety", "traffic", "social-dilemma"),
     dict(publication_role="safety-critical MARL benchmark",
          status="needs PettingZoo ParallelEnv wrapper and safety metrics")),
    ("minecraft", "Minecraft-Inspired Crafting World", "open-ended-games", "pettingzoo", "frontier",
     "hierarchical-rl-plus-llm", 10_000_000, 20, 1, None, ("crafting", "open-ended", "llm-agents", "multi-agent"),
     dict(publication_role="open-ended agentic-game case study",
          status="needs PettingZoo wrapper, scripted baselines, and LLM ablation harness")),
    ("simcity", "SimCity Collaborative Construction", "open-ended-games", "pettingzoo", "frontier",
     "hierarchical-rl-plus-llm", 10_000_000, 20, 1, None, ("collaboration", "llm-agents", "economy", "multi-agent"),
     dict(publication_role="LLM/RL collaboration benchmark",
          status="needs PettingZoo wrapper and reproducible LLM transcript evaluation")),
    ("fish", "Fish Schooling", "multi-agent", "pettingzoo", "roadmap", "ippo", 3_000_000, 20, 1, None,
     ("swarm", "predator-prey", "multi-agent"),
     dict(publication_role="swarm behavior benchmark", status="needs PettingZoo wrapper and population-level metrics")),
    ("self-driving-car", "Self-Driving Car Routing", "safety", "pettingzoo", "frontier", "mappo", 5_000_000, 20, 1, None,
     ("traffic", "interpretability", "safety", "multi-agent"),
     dict(publication_role="interpretable safety case study",
          status="needs PettingZoo wrapper, scenario splits, and safety/regret metrics")),
]


def _build() -> dict[str, TaskSpec]:
    tasks = {}
    for tid, title, family, iface, tier, algo, steps, eval_eps, n_envs, thr, tags, extra in _TABLE:
        tasks[tid] = TaskSpec(
            id=tid, title=title, family=family, interface=iface, research_tier=tier, default_algorithm=algo,
            policy_prefix=f"{tid.replace('-', '_')}_policy", total_timesteps=steps, eval_episodes=eval_eps,
            n_envs=n_envs, reward_threshold=thr, tags=tags, **extra)
    return tasks


TASKS: dict[str, TaskSpec] = _build()
CUDA_TASKS = tuple(t for t, s in TASKS.items() if s.trainable)

_ALIASES = {"brick-break": "brickbreak", "food-collector": "foodcollector", "self_driving_car": "self-driving-car"}


def list_tasks(*, include_roadmap: bool = True) -> list[TaskSpec]:
    tasks = [t for t in TASKS.values() if include_roadmap or t.trainable]
    return sorted(tasks, key=lambda t: (t.family, t.id))


def list_task_cards(*, include_roadmap: bool = True) -> list[dict[str, Any]]:
    return [t.card() for t in list_tasks(include_roadmap=include_roadmap)]


def get_task(task_id: str) -> TaskSpec:
    norm = task_id.lower().replace("_", "-")
    key = _ALIASES.get(norm, norm)
    if key not in TASKS:
        raise KeyError(f"Unknown task '{task_id}'. Available: {', '.join(sorted(TASKS))}")
    return TASKS[key]


def make_env(task_id: str):
    task = get_task(task_id)
    if not task.trainable or task.env_factory is None:
        raise ValueError(f"Task '{task_id}' is not a Gymnasium/SB3 trainable task yet.")
    return task.env_factory()
