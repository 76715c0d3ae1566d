"""The reference-shaped Python surface (st / st_cy / prediction / control / dqn / ddpg / merge_gym) on the EMULATED library
with CPU tensors: the very test bodies of tests/test_api_gpu.py, run without a GPU (tests/emu, fixture `emulated_engine`).

What this adds to the CPU suite: the Python host layer -- single-state drop-in calls against the reference's golden vectors,
the batched RL-proposes / MPC-vetoes decision, the batched environment -- is exercised end to end on every change, not only
on the GPU box.  The device run of test_api_gpu.py remains the gate for the product path."""
import pytest

from tests import test_api_gpu as G


@pytest.fixture
def api(emulated_engine):
    import torch
    from rl_mpc_lanemerging_b200 import config, control, ddpg, dqn, merge_gym, prediction, st, st_cy
    config.Settings.CRASH_MIN_S = 20            # configs/st_moderate.json
    config.Settings.OTHER_CAR_SPEED = 11.0
    config.Settings.TEST_ST_STRICTLY_BETTER = False
    return dict(torch=torch, Settings=config.Settings, control=control, ddpg=ddpg, dqn=dqn, merge_gym=merge_gym,
                prediction=prediction, st=st, st_cy=st_cy, device="cpu")


def test_st_dropin_matches_reference_golden(api):
    G.test_st_dropin_matches_reference_golden(api)


def test_st_cy_signature_dropin(api, oracle):
    G.test_st_cy_signature_dropin(api, oracle)


def test_rollout_helpers_match_golden(api):
    G.test_rollout_helpers_match_golden(api)


def test_combined_control_matches_cpu_chain(api, oracle):
    G.test_combined_control_matches_cpu_chain(api, oracle)


def test_merge_env_steps(api):
    G.test_merge_env_steps(api)


@pytest.mark.parametrize("fused", [False, True])
def test_env_step_follows_the_reference_tick(api, fused):
    G.test_env_step_follows_the_reference_tick(api, fused)


def test_fused_rollout_step_equals_its_pieces(api):
    G.test_fused_rollout_step_equals_its_pieces(api)


def test_predict_step_without_ego_method(api, oracle):
    from tests import test_without_ego_gpu as W
    W.test_highway_state_method(oracle)


def test_published_actor_matches_the_reference_forward(api):
    G.test_published_actor_matches_the_reference_forward(api)


@pytest.mark.parametrize("tag", ["plain", "b"])
def test_combined_control_matches_the_reference(api, tag):
    G.test_combined_control_matches_the_reference(api, tag)


def test_legacy_pickle_loader_on_the_published_checkpoint():
    """DDPGAgent.load's unpickling of the reference's whole-module checkpoints (ddpg.py:37-44), on the real file where the reference
    tree is mounted (this container): the tensors are the committed fixture's."""
    import os
    import numpy as np
    from rl_mpc_lanemerging_b200 import ddpg
    path = "/root/reference/pretrained_models/ddpg_moderate1_extended/policy.pt"
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted (GPU box)")
    sd = ddpg._load_legacy_state_dict(path)
    P = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "policy_moderate1.npz"))
    assert sorted(sd) == sorted(f"model.{i}.{w}" for i in (0, 2, 4) for w in ("weight", "bias"))
    for k, v in sd.items():
        assert np.array_equal(v.numpy(), P[k.replace(".", "_")]), k
