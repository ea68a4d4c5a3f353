"""Entity / agent data model and the batched World (reference: multiagent/core.py).

The reference's unit of work is one world made of Python objects holding 2-vectors in float64
(core.py:4-79).  Here a `World` describes the *topology and properties* of a world exactly the
same way (lists of `Agent` / `Landmark` objects with `size`, `movable`, `collide`, `accel`,
`max_speed`, ... attributes that `Scenario.make_world()` fills in), while the *state* of
`num_envs` independent copies of that world lives in struct-of-arrays fp32 CUDA tensors owned by
a `NativeWorld` (native.py).  `entity.state.p_pos` / `p_vel` / `c` and `agent.action.u` / `c`
are properties that read and write those tensors:

  * scalar mode  (`world.num_envs is None`, the reference-compatible default): getters return
    float64 NumPy copies of world 0 with the reference's shapes, e.g. `p_pos.shape == (2,)`;
  * batched mode (`world.num_envs = N`): getters return live CUDA tensor views `[N, 2]`.

`World.step()` (core.py:117-131) launches the sm_100a physics kernel.  There is no CPU code path.
"""
import numpy as np

from . import _lib

_PROGRAM_IDS = {
    "simple": _lib.SCN_SIMPLE,
    "simple_spread": _lib.SCN_SPREAD,
    "simple_tag": _lib.SCN_TAG,
    "simple_world_comm": _lib.SCN_WORLD_COMM,
    "simple_adversary": _lib.SCN_ADVERSARY,
    "simple_push": _lib.SCN_PUSH,
    "simple_speaker_listener": _lib.SCN_SPEAKER_LISTENER,
    "simple_reference": _lib.SCN_REFERENCE,
    "simple_crypto": _lib.SCN_CRYPTO,
    "custom": _lib.SCN_CUSTOM,     # user scenario: native physics, observation / reward in the user's torch code
}


class EntityState(object):
    """physical state of an entity (core.py:4-9)"""

    def __init__(self):
        self._world = None
        self._kind = None
        self._index = -1
        self._local = {}

    def _attach(self, world, kind, index):
        self._world, self._kind, self._index = world, kind, index

    def _get(self, field):
        w = self._world
        if w is not None and w._native is not None:
            return w._read_state(self._kind, self._index, field)
        return self._local.get(field)

    def _set(self, field, value):
        w = self._world
        if w is not None and w._native is not None:
            w._write_state(self._kind, self._index, field, value)
        else:
            self._local[field] = value

    p_pos = property(lambda self: self._get("p_pos"), lambda self, v: self._set("p_pos", v))
    p_vel = property(lambda self: self._get("p_vel"), lambda self, v: self._set("p_vel", v))


class AgentState(EntityState):
    """adds the communication utterance (core.py:11-16)"""
    c = property(lambda self: self._get("c"), lambda self, v: self._set("c", v))


class Action(object):
    """physical action u and communication action c of an agent (core.py:19-24)"""

    def __init__(self):
        self._world = None
        self._index = -1
        self._local = {}

    def _attach(self, world, index):
        self._world, self._index = world, index

    def _get(self, field):
        w = self._world
        if w is not None and w._native is not None:
            return w._read_action(self._index, field)
        return self._local.get(field)

    def _set(self, field, value):
        w = self._world
        if w is not None and w._native is not None:
            w._write_action(self._index, field, value)
        else:
            self._local[field] = value

    u = property(lambda self: self._get("u"), lambda self, v: self._set("u", v))
    c = property(lambda self: self._get("c"), lambda self, v: self._set("c", v))


class Entity(object):
    """properties of a physical world entity; defaults as in core.py:27-51"""

    def __init__(self):
        self.name = ''
        self.size = 0.050
        self.movable = False
        self.collide = True
        self.density = 25.0
        self.color = None
        self.max_speed = None
        self.accel = None
        self.state = EntityState()
        self.initial_mass = 1.0

    @property
    def mass(self):
        return self.initial_mass


class Landmark(Entity):
    def __init__(self):
        super(Landmark, self).__init__()


class Agent(Entity):
    """agent defaults as in core.py:54-79"""

    def __init__(self):
        super(Agent, self).__init__()
        self.movable = True
        self.silent = False
        self.blind = False
        self.u_noise = None
        self.c_noise = None
        self.u_range = 1.0
        self.state = AgentState()
        self.action = Action()
        self.action_callback = None


class World(object):
    """A batch of identical-topology particle worlds (reference: core.py:82-196)."""

    def __init__(self, num_envs=None, device=None):
        self.agents = []
        self.landmarks = []
        self.dim_c = 0
        self.dim_p = 2
        self.dim_color = 3
        self.dt = 0.1
        self.damping = 0.25
        self.contact_force = 1e+2
        self.contact_margin = 1e-3
        # ---- batch binding (not in the reference) ----
        self.num_envs = num_envs          # None -> scalar, reference-compatible API
        self.device = device
        self.seed = 0
        self.world_offset = 0             # global index of this shard's first world
        self.native_program = None        # name of the compiled scenario program
        self.scenario = None
        self._native = None
        self._shape_handle = None
        self._obs_valid = False
        self._needs_reset = True

    # ---- topology (core.py:101-115) --------------------------------------------------------
    @property
    def entities(self):
        return self.agents + self.landmarks

    @property
    def policy_agents(self):
        return [agent for agent in self.agents if agent.action_callback is None]

    @property
    def scripted_agents(self):
        return [agent for agent in self.agents if agent.action_callback is not None]

    @property
    def batched(self):
        return self.num_envs is not None

    @property
    def batch_size(self):
        return 1 if self.num_envs is None else int(self.num_envs)

    # ---- descriptor --------------------------------------------------------------------------
    def descriptor(self):
        """Flatten what make_world() wrote onto this object into the C-ABI `mpe_desc`."""
        if self.native_program not in _PROGRAM_IDS:
            raise NotImplementedError(
                "World has no native sm_100a scenario program (world.native_program=%r). The built-in scenarios "
                "are compiled; a user scenario must derive from TorchScenario (native physics, observation / "
                "reward written with torch ops on the device). There is deliberately no CPU fallback."
                % (self.native_program,))
        if self.scripted_agents:
            raise NotImplementedError("scripted agents (action_callback) are not supported by the native path")
        if self.dim_p != 2:
            raise NotImplementedError("dim_p must be 2")
        A, L = len(self.agents), len(self.landmarks)
        if not (1 <= A <= _lib.MPE_MAX_AGENTS and 0 <= L <= _lib.MPE_MAX_LANDMARKS):
            raise ValueError("unsupported entity counts: %d agents, %d landmarks" % (A, L))
        d = _lib.MpeDesc()
        d.abi_version = _lib.MPE_ABI_VERSION
        d.scenario = _PROGRAM_IDS[self.native_program]
        d.n_agents, d.n_landmarks, d.dim_c = A, L, int(self.dim_c)
        d.dt, d.damping = float(self.dt), float(self.damping)
        d.contact_force, d.contact_margin = float(self.contact_force), float(self.contact_margin)
        n_adv = 0
        for i, ag in enumerate(self.agents):
            if ag.u_noise or ag.c_noise:
                raise NotImplementedError("u_noise / c_noise are None in every reference scenario; not supported")
            d.agent_size[i] = float(ag.size)
            d.agent_mass[i] = float(ag.mass)
            d.agent_sens[i] = float(ag.accel) if ag.accel is not None else 5.0   # environment.py:178-181
            d.agent_max_speed[i] = float(ag.max_speed) if ag.max_speed is not None else -1.0
            d.agent_movable[i] = 1 if ag.movable else 0
            d.agent_collide[i] = 1 if ag.collide else 0
            d.agent_silent[i] = 1 if ag.silent else 0
            adv = bool(getattr(ag, "adversary", False))
            d.agent_adversary[i] = 1 if adv else 0
            d.agent_leader[i] = 1 if getattr(ag, "leader", False) else 0
            n_adv += adv
        d.n_adversaries = n_adv
        for l, lm in enumerate(self.landmarks):
            if lm.movable:
                raise NotImplementedError("movable landmarks do not occur in the reference scenarios; not supported")
            d.landmark_size[l] = float(lm.size)
            d.landmark_collide[l] = 1 if lm.collide else 0
        food, forests = getattr(self, "food", []), getattr(self, "forests", [])
        d.n_food, d.n_forests = len(food), len(forests)
        d.n_obstacles = L - len(food) - len(forests)
        if food or forests:  # simple_world_comm.py:52-53: landmarks = obstacles ++ food ++ forests
            if list(self.landmarks[d.n_obstacles:]) != list(food) + list(forests):
                raise ValueError("world.landmarks must be obstacles ++ food ++ forests")
        return d

    def native_shapes(self):
        """Device-less library handle: answers shape queries on machines without a GPU."""
        if self._native is not None:
            return self._native
        if self._shape_handle is None:
            from .native import ShapeHandle
            self._shape_handle = ShapeHandle(self.descriptor(), self.batch_size)
        return self._shape_handle

    # ---- binding -----------------------------------------------------------------------------
    def bind(self, num_envs=None, device=None):
        """Allocate the batch state on the device and attach entity state/action properties."""
        if num_envs is not None:
            self.num_envs = num_envs
        if device is not None:
            self.device = device
        if self._native is not None:
            return self._native
        from .native import NativeWorld
        native = NativeWorld(self.descriptor(), self.batch_size, self.device, seed=self.seed,
                             world_offset=self.world_offset)
        pending = []
        for i, ag in enumerate(self.agents):
            ag.state._attach(self, "agent", i)
            ag.action._attach(self, i)
            pending += [(ag.state, k, v) for k, v in ag.state._local.items() if v is not None]
            pending += [(ag.action, k, v) for k, v in ag.action._local.items() if v is not None]
        for l, lm in enumerate(self.landmarks):
            lm.state._attach(self, "landmark", l)
            pending += [(lm.state, k, v) for k, v in lm.state._local.items() if v is not None and k == "p_pos"]
        self._native = native
        if self._needs_reset:
            self.reset_states()
        for obj, k, v in pending:   # values assigned before binding win over the initial reset
            obj._set(k, v)
            obj._local.pop(k, None)
        return native

    @property
    def native(self):
        return self.bind()

    # ---- World.step (core.py:117-131) ---------------------------------------------------------
    def step(self):
        nw = self.bind()
        nw.world_step()
        self._obs_valid = False

    # ---- reset ---------------------------------------------------------------------------------
    def reset_states(self, mask=None, seed=None):
        nw = self.bind() if self._native is None else self._native
        self._needs_reset = False
        if seed is not None:     # (seed, epoch) key the Philox draw: reseeding restarts the epoch, so that
            self.seed = seed     # env.reset(seed=s) reproduces the same episode every time it is called
            nw.seed = seed
            nw.epoch = 0
            if nw._epoch_dev is not None:
                nw._epoch_dev.zero_()
        nw.reset(mask)
        self._obs_valid = False

    # ---- state access used by the entity properties -------------------------------------------
    def _scalar(self, t):
        return t[0].detach().to("cpu").numpy().astype(np.float64)

    def _read_state(self, kind, index, field):
        import torch
        nw = self._native
        if kind == "agent":
            if field == "p_pos":
                t = nw.agent_pv[index, :, 0:2]
            elif field == "p_vel":
                t = nw.agent_pv[index, :, 2:4]
            else:
                s = nw.speaker_slot(index)
                if s < 0:
                    t = torch.zeros(nw.n_env, nw.dim_c, device=nw.device)
                else:
                    t = nw.comm[s * nw.dim_c:(s + 1) * nw.dim_c, :].t()
        else:
            if field == "p_pos":
                t = nw.lm_p[index]
            else:
                t = torch.zeros(nw.n_env, 2, device=nw.device)
        if self.batched:
            self._obs_valid = False  # the caller holds a writable view
            return t
        return self._scalar(t)

    def _as_tensor(self, value, width):
        import torch
        nw = self._native
        t = torch.as_tensor(np.asarray(value, dtype=np.float32) if not torch.is_tensor(value) else value,
                            dtype=torch.float32, device=nw.device)
        return t.reshape(-1, width) if t.dim() <= 1 else t

    def _write_state(self, kind, index, field, value):
        nw = self._native
        self._obs_valid = False
        if value is None:
            return
        if kind == "agent":
            if field == "p_pos":
                nw.agent_pv[index, :, 0:2] = self._as_tensor(value, 2)
            elif field == "p_vel":
                nw.agent_pv[index, :, 2:4] = self._as_tensor(value, 2)
            else:
                s = nw.speaker_slot(index)
                if s >= 0 and nw.dim_c > 0:
                    nw.comm[s * nw.dim_c:(s + 1) * nw.dim_c, :] = self._as_tensor(value, nw.dim_c).t()
        elif field == "p_pos":
            nw.lm_p[index] = self._as_tensor(value, 2)
        # landmark velocity is identically zero (no scenario has a movable landmark)

    def _read_action(self, index, field):
        import torch
        nw = self._native
        if field == "u":
            t = nw.act_u[index]
        else:
            s = nw.speaker_slot(index)
            t = torch.zeros(nw.n_env, nw.dim_c, device=nw.device) if s < 0 else \
                nw.act_c[s * nw.dim_c:(s + 1) * nw.dim_c, :].t()
        return t if self.batched else self._scalar(t)

    def _write_action(self, index, field, value):
        nw = self._native
        if value is None:
            return
        if field == "u":
            nw.act_u[index] = self._as_tensor(value, 2)
        else:
            s = nw.speaker_slot(index)
            if s >= 0 and nw.dim_c > 0:
                nw.act_c[s * nw.dim_c:(s + 1) * nw.dim_c, :] = self._as_tensor(value, nw.dim_c).t()

    # ---- scenario callbacks: slices of the native observe kernel's outputs ---------------------
    def _observe_if_stale(self):
        """scenario.observation / reward / benchmark_data called directly (outside env.step): the observe kernel
        runs into a slab of its OWN, never into the buffers `step` / `reset` hand out (with env.reuse_buffers those
        alias the caller's tensors).  Per-agent rewards, no shared-reward sum: that is MultiAgentEnv.step's glue."""
        nw = self.bind()
        if nw.cb_out is None:
            nw.cb_out = nw.persistent_outputs()
        if not self._obs_valid:
            nw.observe(nw.cb_out, flags=0)
            self._obs_valid = not self.batched  # batched getters hand out writable views
        return nw

    def _agent_index(self, agent):
        for i, a in enumerate(self.agents):
            if a is agent:
                return i
        raise ValueError("agent does not belong to this world")

    def native_observation(self, agent):
        nw = self._observe_if_stale()
        o = nw.cb_out.obs[self._agent_index(agent)]
        return o if self.batched else self._scalar(o)

    def native_reward(self, agent):
        nw = self._observe_if_stale()
        r = nw.cb_out.rew[self._agent_index(agent)]
        return r if self.batched else float(r[0].item())

    def native_benchmark_data(self, agent):
        nw = self._observe_if_stale()
        return nw.benchmark_data(self._agent_index(agent), self.batched, nw.cb_out)
