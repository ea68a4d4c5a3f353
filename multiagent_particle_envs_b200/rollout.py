"""On-device rollouts: policy -> env.step -> policy -> ... captured once in a CUDA graph and replayed
(removes the per-step launch and Python overhead that dominates below ~100k worlds).

    rollout = GraphedRollout(env, policy, steps=25)      # policy(obs_n) -> action_n, all CUDA tensors
    obs_T, rew_sum = rollout.run()                       # one graph launch = `steps` fused env steps

`policy` must be capturable (pure torch CUDA ops, static shapes).  The environment writes into its
persistent output slab (`env.reuse_buffers`), so nothing is allocated while the graph runs.
"""


class GraphedRollout(object):
    def __init__(self, env, policy, steps, warmup=2, reset_every=None):
        import torch
        if not env.world.batched:
            raise ValueError("GraphedRollout needs a batched env (make_env(..., num_envs=N))")
        self.env, self.policy, self.steps = env, policy, int(steps)
        self.reset_every = reset_every        # e.g. 25: env.reset() inside the graph every 25 steps (fresh draws per replay)
        self.torch = torch
        env.reuse_buffers = True
        nw = env.world.bind()
        nw.enable_device_epoch()
        self.obs = [o.clone() for o in env.reset()]          # static input buffers of the graph
        self.rew_sum = torch.zeros(env.n, nw.n_env, device=nw.device)
        self.stream = torch.cuda.Stream(nw.device)
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):                          # warm-up outside capture (lazy inits, autotune)
                self._body()
            self.stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self._body()

    def _body(self):
        torch = self.torch
        obs = self.obs
        self.rew_sum.zero_()
        for t in range(self.steps):
            with torch.no_grad():
                act = self.policy(obs)
            obs, rew_n, done_n, _ = self.env.step([a.contiguous() for a in act])
            self.rew_sum += torch.stack(list(rew_n))
            if self.reset_every and (t + 1) % self.reset_every == 0:
                obs = self.env.reset()
        for dst, src in zip(self.obs, obs):                  # the next replay continues from here
            dst.copy_(src)

    def run(self):
        self.graph.replay()
        return self.obs, self.rew_sum
