"""Learning evidence for the DDPG trainer (reference ddpg.py:46-117): trains in the SUMO-free world with the
reference's two-stage schedule (learning rate / 10 for the second stage) and evaluates the GREEDY actor on fresh episodes after
every chunk of frames.  Writes one JSON object (stdout): merge / crash rate, mean return per episode and mean |jerk| over training.

    python tools/learning_curve.py [frames_per_stage=200000] [chunks_per_stage=4] [eval_episodes=512]

Config: train_moderate_1.json semantics (traffic 11 m/s / 1.2 s, ALT_J_WEIGHT 0.1, SEED 100)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from rl_mpc_lanemerging_b200 import control, ddpg, merge_gym, st
from rl_mpc_lanemerging_b200.config import Settings

frames = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200000
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n_eval = int(sys.argv[3]) if len(sys.argv) > 3 else 512
Settings.reset()
for k, v in dict(TASK="TRAIN_DDPG", BASE_TRAFFIC_INTERVAL=1.2, OTHER_CAR_SPEED=11.0, ALT_J_WEIGHT=0.1, CRASH_MIN_S=20, SEED=100,
                 LEARNING_RATE=2e-4).items():
    setattr(Settings, k, v)
st.refresh_engine()


def evaluate(tr):
    """Greedy actor on n_eval fresh episodes: (merged, crashed, mean return, mean |jerk|)."""
    agent = ddpg.DDPGAgent(device=tr.device)
    agent.policy.load_state_dict(tr.policy.state_dict())
    env = merge_gym.MergeEnv(n_eval, seed=4242, auto_reset=False)
    env.reset()
    ret = torch.zeros(n_eval, dtype=torch.float64, device=env.device)
    alive = torch.ones(n_eval, dtype=torch.bool, device=env.device)
    merged = torch.zeros_like(alive); crashed = torch.zeros_like(alive)
    jerk_sum = torch.zeros_like(ret); steps = torch.zeros_like(ret)
    for _ in range(int(Settings.MAX_EPISODE_LENGTH / Settings.TICK_LENGTH) + 1):
        jerk = agent.get_control(env.state).double()
        _o, r, done, info = env.step(jerk)
        ret += torch.where(alive, r, torch.zeros_like(r))
        jerk_sum += torch.where(alive, info["projected_jerk"].abs(), torch.zeros_like(r)); steps += alive.double()
        merged |= alive & info["merged"]; crashed |= alive & info["crashed"]
        alive &= ~done
        if not bool(alive.any()):
            break
    return dict(merged=float(merged.double().mean()), crashed=float(crashed.double().mean()), mean_return=float(ret.mean()),
                mean_abs_jerk=float((jerk_sum / steps.clamp(min=1)).mean()))


tr = ddpg.DDPGAgent._trainer(seed=100)
curve = [dict(frames=0, stage=0, lr=tr.lr, **evaluate(tr))]
t0 = time.time()
for stage in (1, 2):
    if stage == 2:
        tr.set_lr(tr.lr / 10)                                   # ddpg.py:99: LEARNING_RATE /= 10, training resumes from the same weights
    for c in range(chunks):
        tr.train(frames // chunks)
        curve.append(dict(frames=tr.frames, stage=stage, lr=tr.lr, grad_steps=tr.grad_steps, wall_s=time.time() - t0, **evaluate(tr)))
        print(curve[-1], file=sys.stderr, flush=True)
res = dict(config="train_moderate_1.json semantics; SUMO-free world (Settings.WORLD_MODEL default); greedy evaluation on %d fresh episodes" % n_eval,
           num_envs=tr.env.B, minibatch=tr.h["minibatch_size"], updates_per_tick=tr.h.get("updates_per_tick"), frames_per_stage=frames,
           improved=bool(curve[-1]["mean_return"] > curve[0]["mean_return"]), curve=curve)
print(json.dumps(res, indent=1))
