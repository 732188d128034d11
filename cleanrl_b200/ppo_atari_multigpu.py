"""Drop-in for cleanrl/ppo_atari_multigpu.py: data-parallel PPO, one process per GPU.

    torchrun --standalone --nnodes=1 --nproc_per_node=8 cleanrl_b200/ppo_atari_multigpu.py --backend nccl

Same flags / Agent / tags / rank-aware seeding as the reference
(cleanrl/ppo_atari_multigpu.py:29-102,163-231,387-397).  What changes underneath: all gradients
already live in ONE flat fp32 buffer, so the exchange of ppo_atari_multigpu.py:360-374 (cat ->
all_reduce -> 12 copy-backs with a division) is a single in-place all-reduce of that buffer; the
``/ world_size`` is folded into the fused clip+Adam kernel.
"""
from __future__ import annotations

import os
import random
import sys
import time
import warnings

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from cleanrl_b200 import cli
from cleanrl_b200.agents import NatureCNNAgent as Agent, layer_init  # noqa: F401
from cleanrl_b200.ppo_engine import PPOEngine

Args = cli.ppo_atari_multigpu_args(os.path.basename(__file__)[: -len(".py")])
run_name = None


from cleanrl_b200.ppo_atari import make_env  # noqa: E402,F401  (same wrapper stack, ppo_atari_multigpu.py:105-124)


def make_envs(args, run_name):
    """gymnasium SyncVectorEnv of wrapped Atari envs as the reference (ppo_atari_multigpu.py:225-227) when
    gymnasium + cleanrl_utils are importable; otherwise the seeded synthetic gymnasium-style vec env."""
    if not cli.use_synthetic(args):
        try:
            import gymnasium as gym  # type: ignore  # noqa: F401
        except ImportError as e:
            raise cli.env_import_error("gymnasium (+ ale-py, cleanrl_utils.atari_wrappers)", e) from e
        return gym.vector.SyncVectorEnv([make_env(args.env_id, i, args.capture_video, run_name)
                                         for i in range(args.local_num_envs)])
    from cleanrl_b200.synthetic_envs import SyntheticGymnasiumVec

    return SyntheticGymnasiumVec(args.local_num_envs, kind="atari")


def main(argv=None, writer_factory=None, env_factory=None, on_iteration=None, agent_hook=None):
    global run_name
    args = cli.parse(Args, argv)
    local_rank = int(os.getenv("LOCAL_RANK", "0"))
    args.world_size = int(os.getenv("WORLD_SIZE", "1"))
    args.local_batch_size = int(args.local_num_envs * args.num_steps)
    args.local_minibatch_size = int(args.local_batch_size // args.num_minibatches)
    args.num_envs = args.local_num_envs * args.world_size
    args.batch_size = int(args.num_envs * args.num_steps)
    args.minibatch_size = int(args.batch_size // args.num_minibatches)
    args.num_iterations = args.total_timesteps // args.batch_size
    if args.world_size > 1:
        if not dist.is_initialized():
            dist.init_process_group(args.backend, rank=local_rank, world_size=args.world_size)
    else:
        warnings.warn("Not using distributed mode! Launch with `torchrun --standalone --nnodes=1 "
                      "--nproc_per_node=N cleanrl_b200/ppo_atari_multigpu.py` to use N GPUs.")
    cli.use_synthetic(args)
    run_name = cli.run_name_for(args)
    writer = None
    if local_rank == 0:
        if args.track:
            import wandb

            wandb.init(project=args.wandb_project_name, entity=args.wandb_entity, sync_tensorboard=True,
                       config=vars(args), name=run_name, monitor_gym=True, save_code=True)
        if writer_factory is None:
            from torch.utils.tensorboard import SummaryWriter as writer_factory
        writer = writer_factory(f"runs/{run_name}")
        writer.add_text("hyperparameters",
                        "|param|value|\n|-|-|\n%s" % ("\n".join([f"|{k}|{v}|" for k, v in vars(args).items()])))

    # seeding: a different env / numpy / python stream per rank, the SAME torch stream for the weights
    # (ppo_atari_multigpu.py:207-212), torch re-seeded per rank after the model exists (:231)
    args.seed += local_rank
    random.seed(args.seed)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed - local_rank)
    torch.backends.cudnn.deterministic = args.torch_deterministic

    if not (torch.cuda.is_available() and args.cuda) and not PPOEngine.ALLOW_NON_CUDA_FOR_TESTS:
        raise RuntimeError("cleanrl_b200.ppo_atari_multigpu runs on libb200rl CUDA kernels: CUDA devices and "
                           "--cuda are required (no CPU fallback).")
    if not torch.cuda.is_available():
        device = torch.device("cpu")       # only reachable from the CPU test harness (tests/cpu_backend.py)
    elif len(args.device_ids) > 0:
        assert len(args.device_ids) == args.world_size, \
            "you must specify the same number of device ids as `--nproc_per_node`"
        device = torch.device(f"cuda:{args.device_ids[local_rank]}")
    else:
        device = torch.device(f"cuda:{local_rank}" if torch.cuda.device_count() >= args.world_size else "cuda")
    if device.type == "cuda":
        torch.cuda.set_device(device)

    envs = env_factory(args) if env_factory else make_envs(args, run_name)
    assert hasattr(envs.single_action_space, "n"), "only discrete action space is supported"
    agent = Agent(envs).to(device)
    agent.precision = args.precision
    torch.manual_seed(args.seed)
    if agent_hook:
        agent_hook(agent)
    all_reduce = (lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)) if args.world_size > 1 else None
    eng_args = argparse_like(args)
    engine = PPOEngine(agent, eng_args, envs.single_observation_space.shape, envs.single_observation_space.dtype,
                       args.local_num_envs, device, world_size=args.world_size,
                       gae_mode=0 if args.gae_kernel == "sequential" else 1, all_reduce=all_reduce)

    global_step = 0
    start_time = time.time()
    next_obs, _ = envs.reset(seed=args.seed)
    next_done = np.zeros(args.local_num_envs, dtype=np.float32)
    lrnow = args.learning_rate

    for iteration in range(1, args.num_iterations + 1):
        if args.anneal_lr:
            frac = 1.0 - (iteration - 1.0) / args.num_iterations
            lrnow = frac * args.learning_rate
        action = None
        for step in range(0, args.num_steps):
            global_step += args.num_envs          # counts the GLOBAL envs (ppo_atari_multigpu.py:257)
            action = engine.policy_step(step, next_obs, next_done)
            next_obs, reward, terminations, truncations, infos = envs.step(action)
            next_done = np.logical_or(terminations, truncations)
            engine.record_reward(step, reward)
            if not writer:
                continue
            if "final_info" in infos:
                for info in infos["final_info"]:
                    if info and "episode" in info:
                        print(f"global_step={global_step}, episodic_return={info['episode']['r']}")
                        writer.add_scalar("charts/episodic_return", info["episode"]["r"], global_step)
                        writer.add_scalar("charts/episodic_length", info["episode"]["l"], global_step)
        print(f"local_rank: {local_rank}, action.sum(): {int(np.sum(action))}, iteration: {iteration}, "
              f"agent.actor.weight.sum(): {float(agent.actor.weight.sum())}")

        engine.finish_rollout(next_obs, next_done)
        st = engine.update(lrnow)
        explained_var = engine.explained_variance()

        if local_rank == 0:
            writer.add_scalar("charts/learning_rate", lrnow, global_step)
            writer.add_scalar("losses/value_loss", st["v_loss"], global_step)
            writer.add_scalar("losses/policy_loss", st["pg_loss"], global_step)
            writer.add_scalar("losses/entropy", st["entropy"], global_step)
            writer.add_scalar("losses/old_approx_kl", st["old_approx_kl"], global_step)
            writer.add_scalar("losses/approx_kl", st["approx_kl"], global_step)
            writer.add_scalar("losses/clipfrac", st["clipfrac_mean"], global_step)
            writer.add_scalar("losses/explained_variance", explained_var, global_step)
            sps = int(global_step / (time.time() - start_time))
            print("SPS:", sps)
            writer.add_scalar("charts/SPS", sps, global_step)
        if on_iteration is not None:
            on_iteration(iteration, engine, st)

    envs.close()
    if local_rank == 0:
        writer.close()
        if args.track:
            import wandb

            wandb.finish()
    return engine


def argparse_like(args):
    """The engine reads PPO hyper-parameters by attribute; hand it the parsed dataclass unchanged."""
    return args


if __name__ == "__main__":
    main()
