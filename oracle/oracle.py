"""TEST INFRASTRUCTURE ONLY: ctypes front-end of the CPU oracle (oracle/mpe_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module, and only as the checker / the timed CPU arm -- never on the product path.

Arrays use the oracle's host layout (array-of-structs per world):
    pv [n,A,4]  lm [n,L,2]  comm [n,A,dim_c]  act [n,sum_act]  obs [n,sum_obs]  rew [n,A]
    done [n,A] uint8  info [n,A,info_dim]
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libmpe_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env=dict(os.environ, CC="gcc"))
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


class Oracle(object):
    def __init__(self, desc, dtype="f64"):
        assert dtype in ("f64", "f32")
        self.lib = load()
        self.desc = desc
        self.sfx = "_" + dtype
        self.np_dtype = np.float64 if dtype == "f64" else np.float32
        self.A, self.L, self.C = desc.n_agents, desc.n_landmarks, desc.dim_c
        f = lambda name: getattr(self.lib, name + self.sfx)  # noqa: E731
        self.act_dims = [f("mpe_oracle_act_dim")(ctypes.byref(desc), i) for i in range(self.A)]
        self.obs_dims = [f("mpe_oracle_obs_dim")(ctypes.byref(desc), i) for i in range(self.A)]
        self.info_dim = f("mpe_oracle_info_dim")(ctypes.byref(desc))
        self._f = f

    def _arr(self, a, shape):
        a = np.ascontiguousarray(a, dtype=self.np_dtype)
        assert a.shape == tuple(shape), (a.shape, shape)
        return a

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None

    def empty_outputs(self, n):
        obs = np.zeros((n, sum(self.obs_dims)), self.np_dtype)
        rew = np.zeros((n, self.A), self.np_dtype)
        done = np.ones((n, self.A), np.uint8)
        info = np.zeros((n, self.A, max(self.info_dim, 1)), self.np_dtype)
        return obs, rew, done, info

    def step(self, pv, lm, comm, act, flags=0, goal=None):
        """MultiAgentEnv.step for n worlds; returns (pv', comm', obs, rew, done, info)"""
        n = pv.shape[0]
        pv = self._arr(pv, (n, self.A, 4)).copy()
        lm = self._arr(lm, (n, self.L, 2))
        comm = self._arr(comm, (n, self.A, self.C)).copy()
        width = sum(self.act_dims)
        if flags & 4:   # MPE_FLAG_DISCRETE_ACTION_INPUT: one integer per sub-action
            d = self.desc
            width = sum((1 if d.agent_movable[i] else 0) + (0 if d.agent_silent[i] else 1) for i in range(self.A))
        act = self._arr(act, (n, width))
        obs, rew, done, info = self.empty_outputs(n)
        g = np.ascontiguousarray(goal, np.int32) if goal is not None else None
        fn = self._f("mpe_oracle_step")
        fn.restype = None
        fn(ctypes.byref(self.desc), ctypes.c_int64(n), self._p(pv), self._p(lm), self._p(comm), self._p(g),
           ctypes.c_int(0 if g is None else g.shape[1]), self._p(act), self._p(obs), self._p(rew),
           self._p(done), self._p(info) if self.info_dim > 0 else None, ctypes.c_uint32(flags))
        return pv, comm, obs, rew, done, info[:, :, :self.info_dim]

    def rollout(self, pv, lm, comm, act, steps, flags=0):
        """`steps` env steps in C on the same actions; arrays are updated in place (already in the oracle dtype)"""
        n = pv.shape[0]
        if not hasattr(self, "_ro") or self._ro[0].shape[0] != n:
            self._ro = self.empty_outputs(n)
        obs, rew, done, _ = self._ro
        fn = self._f("mpe_oracle_rollout")
        fn.restype = None
        fn(ctypes.byref(self.desc), ctypes.c_int64(n), self._p(pv), self._p(lm), self._p(comm), self._p(act),
           ctypes.c_int(steps), self._p(obs), self._p(rew), self._p(done), ctypes.c_uint32(flags))
        return obs, rew, done

    def observe(self, pv, lm, comm, flags=0, goal=None):
        n = pv.shape[0]
        pv = self._arr(pv, (n, self.A, 4))
        lm = self._arr(lm, (n, self.L, 2))
        comm = self._arr(comm, (n, self.A, self.C))
        obs, rew, done, info = self.empty_outputs(n)
        g = np.ascontiguousarray(goal, np.int32) if goal is not None else None
        fn = self._f("mpe_oracle_observe")
        fn.restype = None
        fn(ctypes.byref(self.desc), ctypes.c_int64(n), self._p(pv), self._p(lm), self._p(comm), self._p(g),
           ctypes.c_int(0 if g is None else g.shape[1]), self._p(obs), self._p(rew), self._p(done),
           self._p(info) if self.info_dim > 0 else None, ctypes.c_uint32(flags))
        return obs, rew, done, info[:, :, :self.info_dim]

    def set_action(self, act, flags=0):
        n = act.shape[0]
        act = np.ascontiguousarray(act, self.np_dtype)
        u = np.zeros((n, self.A, 2), self.np_dtype)
        c = np.zeros((n, self.A, max(self.C, 1)), self.np_dtype)
        cc = np.zeros((n, self.A * self.C), self.np_dtype)
        fn = self._f("mpe_oracle_set_action")
        fn.restype = None
        fn(ctypes.byref(self.desc), ctypes.c_int64(n), self._p(act), ctypes.c_uint32(flags), self._p(u), self._p(cc))
        if self.C:
            c = cc.reshape(n, self.A, self.C)
        return u, c

    def world_step(self, pv, lm, comm, u, c):
        n = pv.shape[0]
        pv = self._arr(pv, (n, self.A, 4)).copy()
        lm = self._arr(lm, (n, self.L, 2))
        comm = self._arr(comm, (n, self.A, self.C)).copy()
        u = self._arr(u, (n, self.A, 2))
        c = self._arr(np.asarray(c).reshape(n, self.A, self.C) if self.C else np.zeros((n, self.A, 0)), (n, self.A, self.C))
        fn = self._f("mpe_oracle_world_step")
        fn.restype = None
        fn(ctypes.byref(self.desc), ctypes.c_int64(n), self._p(pv), self._p(lm), self._p(comm), self._p(u), self._p(c))
        return pv, comm
