"""Harness-side mirror of the reference's example solver `CompressibleEulerSolver`
(examples/compressible_euler/solver.h:33-102, solver.cu:75-229) on top of the C ABI.

Storage follows t8gpu::MemoryManager (memory_manager.inl:3-106): one device allocation of
(nb_variables * nb_steps + 1) arrays of `capacity` elements, array index = step * nb_variables + variable, volume last.
"""
import ctypes as C

NVAR = 5
# StepList of the examples: Step0..Step3 + Fluxes (examples/compressible_euler/solver.h)
STEP0, STEP1, STEP2, STEP3, FLUXES, NB_STEPS = 0, 1, 2, 3, 4, 5


class EulerSolver:
    """One rank of the unstructured compressible-Euler solver.

    mode = "fused": one tile-plan kernel per RK stage (flux accumulators never reach HBM).
    mode = "unfused": the reference's schedule (face kernel with atomics, then SSP_3RK_step*), through the
                      reference-shaped C-ABI entry points.
    """

    cfl = 0.7  # solver.h:37

    def __init__(self, conn_host, volumes, dtype, device=None, mode="fused", max_level=4, plan=None):
        import numpy as np
        import torch
        from . import Plan, RankTables, conn_to_device
        self.torch = torch
        self.dtype = dtype
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.mode = mode
        self.max_level = max_level
        self.n = int(conn_host["n_local"])
        self.nf = int(conn_host["n_faces"])
        self.nb = int(conn_host["n_bfaces"])
        # 128-byte aligned rows: capacity padded to a multiple of 32 elements
        self.capacity = max(32, (self.n + 31) // 32 * 32)
        self.buffer = torch.zeros((NVAR * NB_STEPS + 1, self.capacity), dtype=dtype, device=self.device)
        vol = (volumes if isinstance(volumes, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(volumes)))
        vol = vol.to(dtype).to(self.device)
        self.buffer[NVAR * NB_STEPS, :self.n] = vol
        self.next, self.prev = STEP0, STEP3
        self.speed = torch.zeros(max(1, self.nf + self.nb), dtype=dtype, device=self.device)
        self.speed_max = torch.zeros(1, dtype=dtype, device=self.device)
        self.dt_dev = torch.zeros(1, dtype=dtype, device=self.device)   # time step kept on the device (adaptive=True)
        self.plan = None
        if mode == "fused":
            self.plan = plan if plan is not None else Plan(conn_host, dtype)   # plan: e.g. Plan.from_device(...)
        else:
            self.conn = conn_to_device({k: conn_host[k] for k in
                                        ("face_neighbors", "face_normals", "face_areas")}, dtype, self.device)
            self.conn.update(n_faces=self.nf, n_bfaces=self.nb, ranks=None, indices=None)
            self._tables = {s: RankTables([self.variables(s)], self.device) for s in range(NB_STEPS)}

    # --- MemoryManager accessors -----------------------------------------------------------------------
    def variables(self, step):
        """get_own_variables(step): list of NVAR 1-D tensors of length n (views)."""
        return [self.buffer[step * NVAR + k, :self.n] for k in range(NVAR)]

    def volume(self):
        return self.buffer[NVAR * NB_STEPS, :self.n]

    def set_state(self, u):
        """u: (5, n) tensor or array -> variables of step `next`."""
        t = self.torch.as_tensor(u).to(self.dtype).to(self.device)
        self.buffer[self.next * NVAR:(self.next + 1) * NVAR, :self.n] = t

    def state(self):
        return self.buffer[self.next * NVAR:(self.next + 1) * NVAR, :self.n]

    # --- CompressibleEulerSolver::iterate (solver.cu:75-175) -------------------------------------------
    def iterate(self, dt, stream=None, adaptive=False, length=None):
        """One RK3 step.  adaptive (fused mode): the step uses the dt stored in self.dt_dev and leaves the next one
        there, min(dt, cfl * length / vmax) with vmax the stage-3 maximum wave speed (compute_timestep,
        solver.cu:213-229) -- computed on the device, no copy to the host, no synchronisation."""
        from . import flux_faces, rk3_stage, timestep
        self.next, self.prev = self.prev, self.next
        prev = self.variables(self.prev)
        s1, s2, nxt = self.variables(STEP1), self.variables(STEP2), self.variables(self.next)
        vol = self.volume()
        if self.mode == "fused":
            dtd = self.dt_dev if adaptive else None
            self.plan.stage(1, prev, None, s1, vol, dt, stream=stream, dt_dev=dtd)
            self.plan.stage(2, s1, prev, s2, vol, dt, stream=stream, dt_dev=dtd)
            self.plan.stage(3, s2, prev, nxt, vol, dt, speed_max=self.speed_max, stream=stream, dt_dev=dtd)
            if adaptive:
                timestep(self.speed_max, self.cfl, 0.5 ** self.max_level if length is None else length, dt,
                         self.dt_dev, stream)
                return 4
            return 3
        fl = self.variables(FLUXES)
        T = self._tables
        flux_faces(self.conn, T[self.prev], T[FLUXES], self.speed, stream)
        rk3_stage(1, prev, None, s1, fl, vol, dt, stream=stream)
        flux_faces(self.conn, T[STEP1], T[FLUXES], self.speed, stream)
        rk3_stage(2, prev, s1, s2, fl, vol, dt, stream=stream)
        flux_faces(self.conn, T[STEP2], T[FLUXES], self.speed, stream)
        rk3_stage(3, prev, s2, nxt, fl, vol, dt, stream=stream)
        return 6

    # --- CompressibleEulerSolver::compute_timestep (solver.cu:213-229) ---------------------------------
    def max_wave_speed(self):
        """Device scalar with the stage-3 maximum of |uHat| + aHat."""
        from . import max_speed
        if self.mode == "fused":
            return self.speed_max
        return max_speed(self.speed, self.speed_max)

    def compute_timestep(self):
        vmax = float(self.max_wave_speed().item())
        ft = self.torch.float32 if self.dtype == self.torch.float32 else self.torch.float64
        half_pow = float(self.torch.tensor(0.5, dtype=ft) ** self.max_level)
        return float(self.torch.tensor(self.cfl, dtype=ft) * self.torch.tensor(half_pow, dtype=ft) /
                     self.torch.tensor(vmax, dtype=ft))


class SubgridEulerSolver:
    """Harness-side mirror of `SubgridCompressibleEulerSolver<Subgrid<4,4,4>>` / `<Subgrid<4,4>>`
    (examples/subgrid/solver.h:31-108, solver.inl:152-266) for one rank.

    Storage follows t8gpu::SubgridMemoryManager (subgrid_memory_manager.inl:3-106): nb_variables * nb_steps arrays of
    capacity * Subgrid::size cells, per-element volumes in a separate vector.
    mode = "unfused": the reference's schedule through the reference-shaped entry points;
    mode = "fused": one kernel per RK stage (3-D only), fluxes never reach HBM.
    """

    def __init__(self, conn_host, volumes, dtype, device=None, mode="unfused"):
        import numpy as np
        import torch
        from . import RankTables, conn_to_device
        self.torch = torch
        self.dtype = dtype
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.mode = mode
        self.dim = int(conn_host["dim"])
        self.S = 64 if self.dim == 3 else 16
        self.ne = int(conn_host["n_local"])
        self.nc = self.ne * self.S
        self.buffer = torch.zeros((NVAR * NB_STEPS, max(self.nc, 32)), dtype=dtype, device=self.device)
        self.vol = torch.as_tensor(np.ascontiguousarray(volumes)).to(dtype).to(self.device)
        self.next, self.prev = STEP0, STEP3
        keys = ("face_neighbors", "face_normals", "face_areas", "level_diff", "offsets")
        self.conn = conn_to_device({k: conn_host[k] for k in keys}, dtype, self.device)
        self.conn.update(dim=self.dim, n_faces=int(conn_host["n_faces"]), n_bfaces=int(conn_host["n_bfaces"]),
                         ranks=None, indices=None)
        self._tables = {s: RankTables([self.variables(s)], self.device) for s in range(NB_STEPS)}
        self.plan = None
        if mode == "fused":
            from . import SubgridPlan
            self.plan = SubgridPlan(conn_host, volumes, dtype)

    def variables(self, step):
        return [self.buffer[step * NVAR + k, :self.nc] for k in range(NVAR)]

    def set_state(self, u):
        t = self.torch.as_tensor(u).to(self.dtype).to(self.device)
        self.buffer[self.next * NVAR:(self.next + 1) * NVAR, :self.nc] = t

    def state(self):
        return self.buffer[self.next * NVAR:(self.next + 1) * NVAR, :self.nc]

    def iterate(self, dt, stream=None):
        from . import rk3_stage, subgrid_boundary_flux, subgrid_inner_flux, subgrid_outer_flux
        self.next, self.prev = self.prev, self.next
        prev = self.variables(self.prev)
        s1, s2, nxt = self.variables(STEP1), self.variables(STEP2), self.variables(self.next)
        if self.mode == "fused":
            self.plan.stage(1, prev, None, s1, self.vol, dt, stream=stream)
            self.plan.stage(2, s1, prev, s2, self.vol, dt, stream=stream)
            self.plan.stage(3, s2, prev, nxt, self.vol, dt, stream=stream)
            return 3
        fl = self.variables(FLUXES)
        T = self._tables
        launches = 0
        for stage, sin, vin, vout in ((1, self.prev, prev, s1), (2, STEP1, s1, s2), (3, STEP2, s2, nxt)):
            subgrid_inner_flux(self.dim, self.vol, vin, fl, stream)
            if self.conn["n_bfaces"] > 0:
                subgrid_boundary_flux(self.conn, T[sin], T[FLUXES], stream)
                launches += 1
            subgrid_outer_flux(self.conn, T[sin], T[FLUXES], stream)
            rk3_stage(stage, prev, vin if stage > 1 else None, vout, fl, self.vol, dt, cells_per_vol=self.S,
                      stream=stream)
            launches += 3
        return launches
