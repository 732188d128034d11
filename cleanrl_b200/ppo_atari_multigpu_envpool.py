"""Data-parallel PPO over envpool: the script the reference defers ("we could do something like
`ppo_atari_multigpu_envpool.py` to possibly obtain the fastest PPO + Atari possible ... we may need numba to pin the
threads envpool is using in each subprocess", docs/rl-algorithms/ppo.md:1020).

    torchrun --standalone --nnodes=1 --nproc_per_node=8 cleanrl_b200/ppo_atari_multigpu_envpool.py --local-num-envs 1024

One process per GPU.  Each rank owns an envpool (gym API, ``RecordEpisodeStatistics`` as cleanrl/ppo_atari_envpool.py:83-114)
of ``--local-num-envs`` envs and a ``PPOEngine`` over its own rollout; ranks exchange ONE flat gradient per update
(cleanrl/ppo_atari_multigpu.py:360-374), overlapped with the conv backward.  Seeding, ``global_step`` accounting,
rank-0 logging and the per-iteration debug print are the multi-GPU script's (ppo_atari_multigpu.py:207-231,257,284-286).

Thread pinning: before its env pool is created a rank restricts itself to a contiguous slice of the host cores
(``os.sched_setaffinity``; threads spawned afterwards -- envpool's workers -- inherit the mask) and sizes the pool to
that slice, so eight pools do not migrate across each other's cores.
"""
from __future__ import annotations

import os
import random
import sys
import time
import warnings
from collections import deque

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from cleanrl_b200 import cli
from cleanrl_b200.agents import NatureCNNAgent as Agent, layer_init  # noqa: F401
from cleanrl_b200.ppo_atari_envpool import RecordEpisodeStatistics  # noqa: F401  (reference module-level name)
from cleanrl_b200.ppo_engine import PPOEngine

Args = cli.ppo_atari_multigpu_envpool_args(os.path.basename(__file__)[: -len(".py")])
run_name = None


def core_slice(local_rank, world_size, cores=None):
    """The host cores of rank ``local_rank``: a contiguous, equal share of the cores this process may run on."""
    if cores is None:
        cores = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))
    per = max(1, len(cores) // max(1, world_size))
    lo = (local_rank * per) % len(cores)
    return cores[lo:lo + per] or cores[:1]


def pin_rank(local_rank, world_size):
    """Pin this process (and every thread it spawns from now on) to its core slice; returns the slice."""
    mine = core_slice(local_rank, world_size)
    if hasattr(os, "sched_setaffinity"):
        os.sched_setaffinity(0, mine)
    torch.set_num_threads(max(1, min(len(mine), 4)))        # host-side torch ops are tiny here
    return mine


def make_envs(args, num_threads=0, pinned=True):
    """envpool.make(...) as cleanrl/ppo_atari_envpool.py:185-196, seeded per rank; synthetic only on request."""
    envs = None
    if not cli.use_synthetic(args):
        try:
            import envpool  # type: ignore
        except ImportError as e:
            raise cli.env_import_error("envpool", e) from e
        kw = dict(num_threads=num_threads) if num_threads > 0 else {}
        envs = envpool.make(args.env_id, env_type="gym", num_envs=args.local_num_envs, episodic_life=True,
                            reward_clip=True, seed=args.seed, **kw)
    if envs is None:
        from cleanrl_b200.synthetic_envs import SyntheticAtariVec

        envs = SyntheticAtariVec(args.local_num_envs, seed=args.seed, mode=os.environ.get("CLEANRL_B200_SYNTH_MODE", "fresh"),
                                 pinned=pinned)
    envs.num_envs = args.local_num_envs
    envs.single_action_space = envs.action_space
    envs.single_observation_space = envs.observation_space
    envs = RecordEpisodeStatistics(envs)
    assert hasattr(envs.action_space, "n"), "only discrete action space is supported"
    return envs


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
    cores = pin_rank(local_rank, args.world_size) if args.pin_env_threads else core_slice(0, 1)
    env_threads = args.env_threads if args.env_threads > 0 else len(cores)
    if args.world_size > 1:
        if not dist.is_initialized():
            dist.init_process_group(args.backend, rank=local_rank, world_size=args.world_size)
    else:
        warnings.warn("Not using distributed mode! Launch with `torchrun --standalone --nnodes=1 "
                      "--nproc_per_node=N cleanrl_b200/ppo_atari_multigpu_envpool.py` to use N GPUs.")
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

    # seeding as cleanrl/ppo_atari_multigpu.py:207-212,231: per-rank env / numpy / python streams, one torch stream for
    # the initial weights, torch re-seeded per rank once the model exists
    args.seed += local_rank
    random.seed(args.seed)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed - local_rank)
    torch.backends.cudnn.deterministic = args.torch_deterministic

    if not (torch.cuda.is_available() and args.cuda) and not PPOEngine.ALLOW_NON_CUDA_FOR_TESTS:
        raise RuntimeError("cleanrl_b200.ppo_atari_multigpu_envpool runs on libb200rl CUDA kernels: CUDA devices and "
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

    envs = env_factory(args) if env_factory else make_envs(args, num_threads=env_threads)
    agent = Agent(envs).to(device)
    agent.precision = args.precision
    torch.manual_seed(args.seed)
    if agent_hook:
        agent_hook(agent)
    all_reduce = (lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)) if args.world_size > 1 else None
    engine = PPOEngine(agent, args, envs.single_observation_space.shape, envs.single_observation_space.dtype,
                       args.local_num_envs, device, world_size=args.world_size,
                       gae_mode=0 if args.gae_kernel == "sequential" else 1, all_reduce=all_reduce)
    engine.env_cores = cores
    avg_returns = deque(maxlen=20)

    global_step = 0
    start_time = time.time()
    next_obs = envs.reset()
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
            next_obs, reward, next_done, info = envs.step(action)
            engine.record_reward(step, reward)
            if not writer:
                continue
            finished = np.nonzero(np.logical_and(next_done, info["lives"] == 0))[0]
            for idx in finished:
                print(f"global_step={global_step}, episodic_return={info['r'][idx]}")
                avg_returns.append(info["r"][idx])
                writer.add_scalar("charts/avg_episodic_return", np.average(avg_returns), global_step)
                writer.add_scalar("charts/episodic_return", info["r"][idx], global_step)
                writer.add_scalar("charts/episodic_length", info["l"][idx], global_step)
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


if __name__ == "__main__":
    main()
