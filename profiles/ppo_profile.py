"""Profiling driver: one ball3d PPO rollout (T steps) + a few minibatches at BASELINE config-3 sizes.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python profiles/ppo_profile.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from three_mlagents_b200.ppo import CudaPPO
from three_mlagents_b200.vec_env import CudaVecEnv

T = int(os.environ.get("T", "4"))
MB = int(os.environ.get("MB", "2"))
env = CudaVecEnv("ball3d", 65536, seed=1)
# batch_size = one BASELINE minibatch (262144 rows); total = 65536*T rows -> T/4 minibatches per epoch
model = CudaPPO("MlpPolicy", env, seed=1, n_steps=T, batch_size=262144, n_epochs=MB, ent_coef=0.01,
                mlp_impl=os.environ.get("IMPL", "bf16"))
model.collect_rollouts()
model.train()
torch.cuda.synchronize()
print("ok")
