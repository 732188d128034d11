"""Drop-in for cleanrl/dqn_atari.py (config 5 of BASELINE.json: replay sample + TD update) on libb200rl.

Same flags, ``QNetwork`` surface / state_dict keys, epsilon schedule, TensorBoard tags and stdout as the reference
(cleanrl/dqn_atari.py:27-80,108-125,186-242).  The numpy ``ReplayBuffer`` becomes a device-resident uint8 ring
(``cleanrl_b200.replay.DeviceReplayRing``) sampled with the same numpy index stream; the sampled frames are
gathered inside the conv kernels; the TD target / MSE loss / dL/dQ is one kernel; Adam is the fused flat step.
"""
from __future__ import annotations

import os
import random
import sys
import time

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from cleanrl_b200 import cli, ops
from cleanrl_b200.agents import QNetworkAgent as QNetwork, dqn_sync_target, dqn_update
from cleanrl_b200.replay import DeviceReplayRing

Args = cli.dqn_atari_args(os.path.basename(__file__)[: -len(".py")])
run_name = None


def linear_schedule(start_e: float, end_e: float, duration: int, t: int):
    slope = (end_e - start_e) / duration
    return max(slope * t + start_e, end_e)


def make_env(env_id, seed, idx, capture_video, run_name):
    """The reference's wrapped Atari env thunk (dqn_atari.py:83-104; also what dqn_eval.py calls); only usable when
    gymnasium + ALE are installed."""
    def thunk():
        import gymnasium as gym  # type: ignore
        from cleanrl_b200.ppo_atari import make_env as atari_env

        env = atari_env(env_id, idx, capture_video, run_name)()      # same wrapper stack as the PPO Atari scripts
        env.action_space.seed(seed)
        return env

    return thunk


def make_envs(args, run_name):
    """gymnasium Atari SyncVectorEnv as the reference (dqn_atari.py:163-166) when available, else synthetic."""
    if not cli.use_synthetic(args):
        try:
            import gymnasium as gym  # type: ignore  # noqa: F401
        except ImportError as e:
            raise cli.env_import_error("gymnasium (+ ale-py, cleanrl_utils.atari_wrappers)", e) from e
        return gym.vector.SyncVectorEnv([make_env(args.env_id, args.seed + i, i, args.capture_video, run_name)
                                         for i in range(args.num_envs)])
    from cleanrl_b200.synthetic_envs import SyntheticGymnasiumVec

    return SyntheticGymnasiumVec(args.num_envs, kind="atari")


def main(argv=None, writer_factory=None, env_factory=None, on_update=None):
    global run_name
    args = cli.parse(Args, argv)
    assert args.num_envs == 1, "vectorized envs are not supported at the moment"   # dqn_atari.py:135
    cli.use_synthetic(args)
    run_name = cli.run_name_for(args)
    if args.track:
        import wandb

        wandb.init(project=args.wandb_project_name, entity=args.wandb_entity, sync_tensorboard=True,
                   config=vars(args), name=run_name, monitor_gym=True, save_code=True)
    if writer_factory is None:
        from torch.utils.tensorboard import SummaryWriter as writer_factory
    writer = writer_factory(f"runs/{run_name}")
    writer.add_text("hyperparameters",
                    "|param|value|\n|-|-|\n%s" % ("\n".join([f"|{key}|{value}|" for key, value in vars(args).items()])))

    random.seed(args.seed)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    torch.backends.cudnn.deterministic = args.torch_deterministic
    if not (torch.cuda.is_available() and args.cuda):
        raise RuntimeError("cleanrl_b200.dqn_atari runs on libb200rl CUDA kernels: a CUDA device and --cuda are "
                           "required (no CPU fallback).")
    device = torch.device("cuda")

    envs = env_factory(args) if env_factory else make_envs(args, run_name)
    assert hasattr(envs.single_action_space, "n"), "only discrete action space is supported"
    A = int(envs.single_action_space.n)
    q_network = QNetwork(envs).to(device)
    target_network = QNetwork(envs).to(device)
    q_network.precision = target_network.precision = args.precision
    target_network.load_state_dict(q_network.state_dict())
    q_network.flat, target_network.flat
    rb = DeviceReplayRing(args.buffer_size, envs.single_observation_space.shape, args.num_envs, device)
    stats = torch.zeros(2, dtype=torch.float32, device=device)
    start_time = time.time()

    obs, _ = envs.reset(seed=args.seed)
    for global_step in range(args.total_timesteps):
        epsilon = linear_schedule(args.start_e, args.end_e, args.exploration_fraction * args.total_timesteps, global_step)
        if random.random() < epsilon:
            actions = np.array([envs.single_action_space.sample() for _ in range(envs.num_envs)])
        else:
            q_values = q_network.q_values(torch.from_numpy(np.ascontiguousarray(obs)).to(device))
            actions = ops.argmax(q_values).cpu().numpy()

        next_obs, rewards, terminations, truncations, infos = envs.step(actions)
        if "final_info" in infos:
            for info in infos["final_info"]:
                if info and "episode" in info:
                    print(f"global_step={global_step}, episodic_return={info['episode']['r']}")
                    writer.add_scalar("charts/episodic_return", info["episode"]["r"], global_step)
                    writer.add_scalar("charts/episodic_length", info["episode"]["l"], global_step)

        real_next_obs = next_obs.copy()
        for idx, trunc in enumerate(truncations):
            if trunc:
                real_next_obs[idx] = infos["final_observation"][idx]
        rb.add(obs, real_next_obs, actions, rewards, terminations, infos)
        obs = next_obs

        if global_step > args.learning_starts:
            if global_step % args.train_frequency == 0:
                data = rb.sample(args.batch_size)
                dqn_update(q_network, target_network, rb, data, args.gamma, args.learning_rate,
                           huber=args.huber_loss, stats=stats)
                if on_update is not None:
                    on_update(global_step, stats, q_network)
                if global_step % 100 == 0:
                    td_loss, q_mean = stats.cpu().tolist()
                    writer.add_scalar("losses/td_loss", td_loss, global_step)
                    writer.add_scalar("losses/q_values", q_mean, global_step)
                    sps = int(global_step / (time.time() - start_time))
                    print("SPS:", sps)
                    writer.add_scalar("charts/SPS", sps, global_step)
            if global_step % args.target_network_frequency == 0:
                dqn_sync_target(q_network, target_network, args.tau)

    if args.save_model:
        os.makedirs(f"runs/{run_name}", exist_ok=True)
        model_path = f"runs/{run_name}/{args.exp_name}.cleanrl_model"
        torch.save({k: v.detach().cpu() for k, v in q_network.state_dict().items()}, model_path)
        print(f"model saved to {model_path}")
        # evaluation of the saved model as the reference does (dqn_atari.py:248-261): 10 episodes, epsilon = end_e
        from cleanrl_b200.evals import evaluate_q

        eval_args = type(args)(**{**vars(args), "num_envs": 1})
        eval_envs = env_factory(eval_args) if env_factory else make_envs(eval_args, f"{run_name}-eval")
        episodic_returns = evaluate_q(model_path, None, args.env_id, eval_episodes=10, run_name=f"{run_name}-eval",
                                      Model=QNetwork, device=device, epsilon=args.end_e, envs=eval_envs)
        eval_envs.close()
        for idx, episodic_return in enumerate(episodic_returns):
            writer.add_scalar("eval/episodic_return", float(np.asarray(episodic_return).reshape(-1)[0]), idx)
        if args.upload_model:
            print("[cleanrl_b200] --upload-model needs cleanrl_utils.huggingface (not part of the hot path); skipped",
                  file=sys.stderr)

    envs.close()
    writer.close()
    return q_network


if __name__ == "__main__":
    main()
