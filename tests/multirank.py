"""Test helper: P ranks of the unstructured solver emulated on ONE device.

Every rank has its own MemoryManager-layout buffer; ghost reads resolve through (rank, remote index) tables whose
entries point into the other ranks' buffers -- the same code path as peer-mapped buffers on other GPUs.  Stages are
launched rank after rank (no kernel ever waits on another kernel)."""
import numpy as np
import torch

import t8gpu_b200
from t8gpu_b200.solver import FLUXES, NB_STEPS, NVAR, STEP0, STEP1, STEP2, STEP3


class MultiRankEuler:
    def __init__(self, forest, P, dtype, device, mode="fused"):
        self.P, self.dtype, self.device, self.mode = P, dtype, device, mode
        npdt = np.float64 if dtype == torch.float64 else np.float32
        self.off = forest.partition_offsets(P)
        lv, cent, vol, _ = forest.elements()
        self.conn = [forest.connectivity(P, r, dtype=npdt) for r in range(P)]
        self.n = [int(c["n_local"]) for c in self.conn]
        self.buf = []
        for r in range(P):
            cap = max(32, (self.n[r] + 31) // 32 * 32)
            b = torch.zeros((NVAR * NB_STEPS + 1, cap), dtype=dtype, device=device)
            b[NVAR * NB_STEPS, :self.n[r]] = torch.as_tensor(vol[self.off[r]:self.off[r + 1]].astype(npdt)).to(device)
            self.buf.append(b)
        self.tables = {s: t8gpu_b200.RankTables([self.vars(r, s) for r in range(P)], device) for s in range(NB_STEPS)}
        self.next, self.prev = STEP0, STEP3
        if mode == "fused":
            self.plans = [t8gpu_b200.Plan(c, dtype) for c in self.conn]
        else:
            self.dconn = []
            for c in self.conn:
                d = t8gpu_b200.conn_to_device({k: c[k] for k in ("face_neighbors", "face_normals", "face_areas",
                                                                 "ranks", "indices")}, dtype, device)
                d.update(n_faces=c["n_faces"], n_bfaces=c["n_bfaces"])
                self.dconn.append(d)
            self.speed = [torch.zeros(max(1, c["n_faces"] + c["n_bfaces"]), dtype=dtype, device=device)
                          for c in self.conn]

    def vars(self, r, step):
        return [self.buf[r][step * NVAR + k, :self.n[r]] for k in range(NVAR)]

    def vol(self, r):
        return self.buf[r][NVAR * NB_STEPS, :self.n[r]]

    def set_global_state(self, u):
        for r in range(self.P):
            t = torch.as_tensor(np.ascontiguousarray(u[:, self.off[r]:self.off[r + 1]])).to(self.dtype).to(self.device)
            self.buf[r][self.next * NVAR:(self.next + 1) * NVAR, :self.n[r]] = t

    def global_state(self):
        return np.concatenate([self.buf[r][self.next * NVAR:(self.next + 1) * NVAR, :self.n[r]].cpu().numpy()
                               for r in range(self.P)], axis=1)

    def iterate(self, dt):
        self.next, self.prev = self.prev, self.next
        seq = [(1, self.prev, STEP1), (2, STEP1, STEP2), (3, STEP2, self.next)]
        for stage, sin, sout in seq:
            if self.mode == "fused":
                for r in range(self.P):
                    self.plans[r].stage(stage, self.vars(r, sin), self.vars(r, self.prev), self.vars(r, sout),
                                        self.vol(r), dt, in_all=self.tables[sin])
            else:
                for r in range(self.P):  # all fluxes (incl. remote atomics) before any RK update
                    t8gpu_b200.flux_faces(self.dconn[r], self.tables[sin], self.tables[FLUXES], self.speed[r])
                for r in range(self.P):
                    t8gpu_b200.rk3_stage(stage, self.vars(r, self.prev), self.vars(r, sin) if stage > 1 else None,
                                         self.vars(r, sout), self.vars(r, FLUXES), self.vol(r), dt)


class MultiRankSubgrid:
    """P ranks of the fused subgrid solver on ONE device: SubgridMemoryManager-layout buffers per rank, ghost cells read
    through [var][rank] pointer tables, stage s of every rank launched before stage s+1 of any."""

    def __init__(self, forest, P, dtype, device):
        self.P, self.dtype, self.device = P, dtype, device
        npdt = np.float64 if dtype == torch.float64 else np.float32
        self.S = 64 if forest.dim == 3 else 16
        self.off = forest.partition_offsets(P)
        lv, cent, vol, _ = forest.elements()
        self.conn = [forest.connectivity(P, r, subgrid=True, dtype=npdt) for r in range(P)]
        self.n = [int(c["n_local"]) for c in self.conn]
        self.buf, self.vol = [], []
        for r in range(P):
            self.buf.append(torch.zeros((NVAR * NB_STEPS, max(32, self.n[r] * self.S)), dtype=dtype, device=device))
            self.vol.append(torch.as_tensor(vol[self.off[r]:self.off[r + 1]].astype(npdt)).to(device))
        self.tables = {s: t8gpu_b200.RankTables([self.vars(r, s) for r in range(P)], device) for s in range(NB_STEPS)}
        self.plans = [t8gpu_b200.SubgridPlan(self.conn[r], vol[self.off[r]:self.off[r + 1]].astype(npdt), dtype)
                      for r in range(P)]
        self.next, self.prev = STEP0, STEP3

    def vars(self, r, step):
        return [self.buf[r][step * NVAR + k, :self.n[r] * self.S] for k in range(NVAR)]

    def set_global_state(self, u):
        S = self.S
        for r in range(self.P):
            t = torch.as_tensor(np.ascontiguousarray(u[:, self.off[r] * S:self.off[r + 1] * S])).to(self.dtype)
            self.buf[r][self.next * NVAR:(self.next + 1) * NVAR, :self.n[r] * S] = t.to(self.device)

    def global_state(self):
        return np.concatenate([self.buf[r][self.next * NVAR:(self.next + 1) * NVAR, :self.n[r] * self.S].cpu().numpy()
                               for r in range(self.P)], axis=1)

    def iterate(self, dt):
        self.next, self.prev = self.prev, self.next
        for stage, sin, sout in [(1, self.prev, STEP1), (2, STEP1, STEP2), (3, STEP2, self.next)]:
            for r in range(self.P):
                self.plans[r].stage(stage, self.vars(r, sin), self.vars(r, self.prev), self.vars(r, sout), self.vol[r],
                                    dt, in_all=self.tables[sin])
