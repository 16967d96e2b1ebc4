"""DS::centered on a z-decomposed 3-d grid (harness over the C ABI): every rank owns a block of consecutive planes, the field-line
shifts I+ f[k+1] / I- f[k-1] need ONE ghost plane on either side (inc/geometries/mpi_fieldaligned.h:412-424, 476-566: the
reference sends the whole boundary plane to the neighbour in z).  The ghost planes travel with dgb_comm_halo_rows (a "row" is a
plane here), the arithmetic is the single-GPU cell-tiled kernel on the padded block, so every plane is computed with the same
instructions as on one GPU: results are bitwise those of the global call for any number of ranks."""
import ctypes as C
import torch
from ._lib import lib
from ._dev import ptr, stream
from .dist import partition

d = C.c_double


class DistDSCentered:
    def __init__(self, comm, n, Nx, Ny, Nz, plus_csr, minus_csr, bphi_local, delta_phi):
        """plus_csr / minus_csr: (row_offsets, column_indices, values) device tensors of the 2-d matrices I+ / I-;
        bphi_local: this rank's planes of bphi"""
        self.comm = comm
        self.plane = n * n * Nx * Ny
        self.z0, self.planes = partition(Nz, comm.size)[comm.rank]
        self.delta = delta_phi
        self.hp, self.hm = C.c_void_p(), C.c_void_p()
        lib().celltile_plan_create(C.byref(self.hp), n, Nx, Ny, *[ptr(a) for a in plus_csr], stream())
        lib().celltile_plan_create(C.byref(self.hm), n, Nx, Ny, *[ptr(a) for a in minus_csr], stream())
        self.size = self.plane * self.planes
        self.fpad = torch.zeros(self.plane * (self.planes + 2), dtype=torch.float64, device="cuda")
        self.gpad = torch.zeros_like(self.fpad)
        self.bpad = torch.ones_like(self.fpad)
        self.bpad[self.plane:][:self.size].copy_(bphi_local)

    def local(self, v):
        return v[self.z0 * self.plane:(self.z0 + self.planes) * self.plane]

    def centered(self, alpha, f_local, beta, g_local):
        """g = alpha ds_centered(f) + beta g on this rank's planes (ds.h:481-485, 776-786; periodic in z)"""
        self.fpad[self.plane:][:self.size].copy_(f_local)
        self.comm.halo_rows(self.fpad, self.plane, self.planes, 1, True)
        if beta != 0.:
            self.gpad[self.plane:][:self.size].copy_(g_local)
        # planes 1 .. P of the padded block see their true neighbours; the two ghost planes produce values nobody reads
        lib().celltile_ds_centered(self.hp, self.hm, self.planes + 2, d(alpha), ptr(self.fpad), ptr(self.bpad), d(self.delta), d(beta),
                                   ptr(self.gpad), stream())
        g_local.copy_(self.gpad[self.plane:][:self.size])

    def __del__(self):
        try:
            lib().celltile_plan_destroy(self.hp)
            lib().celltile_plan_destroy(self.hm)
        except Exception:
            pass
