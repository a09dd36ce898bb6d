"""Synthetic particle distributions for tests and bench.py.  The reference ships no generator
(InitCond.h:34-48 has them commented out), so these are builder-defined and documented in DESIGN.md.
All values are rounded to fp32 so the fp64 reference and the device path see identical coordinates."""
import numpy as np


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32).astype(np.float64)


def uniform_box(n, seed=12345):
    """cfg 1: x, v ~ U[0,1)^3, m = 1."""
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3), dtype=np.float32)
    vel = rng.random((n, 3), dtype=np.float32)
    return pos.astype(np.float64), vel.astype(np.float64), np.ones(n)


def clustered_small(n, seed=1, nhalo=24, frac=0.5):
    """Small clustered set for exhaustive parity: uniform background + Plummer spheres, unit box, wrapped."""
    rng = np.random.default_rng(seed)
    pos = rng.random((n, 3))
    vel = 0.05 * rng.standard_normal((n, 3))
    nh = int(frac * n)
    centres = rng.random((nhalo, 3))
    which = rng.integers(0, nhalo, nh)
    a = 0.02 * (0.3 + rng.random(nhalo))
    u = rng.random(nh) * 0.99
    r = a[which] / np.sqrt(np.maximum(u, 1e-9) ** (-2.0 / 3.0) - 1.0)
    d = rng.standard_normal((nh, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    pos[:nh] = (centres[which] + r[:, None] * d) % 1.0
    vel[:nh] = 0.2 * rng.standard_normal((nhalo, 3))[which] + 0.03 * rng.standard_normal((nh, 3))
    mass = 1.0 + rng.random(n)
    return _f32(pos), _f32(vel), _f32(mass)


def clustered_box(ng, seed=2025, nhalo=8192, halo_frac=0.25, device="cpu", min_members=64, dims=None, slab=None, return_edges=False):
    """cfg 2-5 generator: a cell-centred lattice of spacing D = 1/ng + Zel'dovich displacement psi (Gaussian field,
    P_psi(k) ~ k^-2, rms |psi| = 1.5 lattice spacings, v = psi), wrapped periodically; a random `halo_frac` of the particles
    is relocated into `nhalo` Plummer spheres:
        centres uniform in the box, membership n_h ~ 1/rank (at least `min_members`), scale radius a_h = 0.3 D (n_h/64)^(1/3),
        r = a / sqrt(u^(-2/3) - 1) with u in (0, 0.99), isotropic directions,
        velocities = bulk N(0, (1.5 D)^2 / 3) per component + N(0, sigma_h^2), sigma_h = 0.3 D (n_h/64)^(1/3).
    dims = (cx, cy, cz): the lattice has cx*ng x cy*ng x cz*ng cells and the periodic box is (cx, cy, cz) long (default (1, 1, 1):
    the ng^3 unit cube).  slab = (rank, world): only the particles whose x falls into slab `rank` of `world` equal slabs along x
    are returned -- every rank of a sharded run generates the SAME global field and keeps its own slab (BASELINE config 5: one
    1024^3 cube cut into slabs); slab = (rank, world, "equal_count") puts the slab faces at the x quantiles instead (equal particle
    counts, the usual load-balanced decomposition); return_edges adds the world + 1 face positions to the result.  Returns float32 torch tensors (pos[N,3], vel[N,3], mass[N]) on `device`; every value is
    therefore exactly representable in fp32 and is widened, not rounded, for the fp64 reference."""
    import torch
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    cx, cy, cz = dims if dims is not None else (1, 1, 1)
    gx, gy, gz = cx * ng, cy * ng, cz * ng
    box = torch.tensor([float(cx), float(cy), float(cz)], device=dev, dtype=torch.float32)
    n = gx * gy * gz
    D = 1.0 / ng
    w = torch.randn((gx, gy, gz), generator=gen, device=dev, dtype=torch.float32)
    wk = torch.fft.rfftn(w)
    del w
    # wave numbers in units of 2 pi / (unit length): n_d / L_d with L_d = c_d
    kx = torch.fft.fftfreq(gx, d=1.0 / gx, device=dev) / cx
    ky = torch.fft.fftfreq(gy, d=1.0 / gy, device=dev) / cy
    kz = torch.fft.rfftfreq(gz, d=1.0 / gz, device=dev) / cz
    k2 = kx[:, None, None] ** 2 + ky[None, :, None] ** 2 + kz[None, None, :] ** 2
    k2[0, 0, 0] = 1.0
    phik = wk / k2
    phik[0, 0, 0] = 0
    del wk, k2
    psi = torch.empty((n, 3), device=dev, dtype=torch.float32)
    for j, kk in enumerate((kx[:, None, None], ky[None, :, None], kz[None, None, :])):
        psi[:, j] = torch.fft.irfftn(1j * kk * phik, s=(gx, gy, gz)).reshape(-1)
    del phik
    rms = torch.sqrt((psi.double() ** 2).sum(1).mean()).item()
    psi *= (1.5 * D / rms)
    ax = [(torch.arange(g, device=dev, dtype=torch.float32) + 0.5) * D for g in (gx, gy, gz)]
    pos = torch.empty((n, 3), device=dev, dtype=torch.float32)
    pos[:, 0] = ax[0][:, None, None].expand(gx, gy, gz).reshape(-1)
    pos[:, 1] = ax[1][None, :, None].expand(gx, gy, gz).reshape(-1)
    pos[:, 2] = ax[2][None, None, :].expand(gx, gy, gz).reshape(-1)
    pos += psi
    vel = psi
    if nhalo > 0 and halo_frac > 0:
        nh_tot = int(halo_frac * n)
        rank = torch.arange(1, nhalo + 1, device=dev, dtype=torch.float64)
        wgt = 1.0 / rank
        memb = torch.clamp((wgt / wgt.sum() * nh_tot).floor().long(), min=min_members)
        # trim / pad the largest halo so that the total is exactly nh_tot
        memb[0] += nh_tot - int(memb.sum().item())
        assert memb[0] > 0, "halo_frac too small for min_members * nhalo"
        hid = torch.repeat_interleave(torch.arange(nhalo, device=dev), memb)
        sel = torch.randperm(n, generator=gen, device=dev)[:nh_tot]
        centres = torch.rand((nhalo, 3), generator=gen, device=dev, dtype=torch.float32) * box
        a = (0.3 * D * (memb.double() / 64.0) ** (1.0 / 3.0)).float()
        sig = (0.3 * D * (memb.double() / 64.0) ** (1.0 / 3.0)).float()
        bulk = torch.randn((nhalo, 3), generator=gen, device=dev, dtype=torch.float32) * (1.5 * D / 3 ** 0.5)
        u = torch.rand(nh_tot, generator=gen, device=dev, dtype=torch.float32) * 0.99
        u = torch.clamp(u, min=1e-6)
        r = a[hid] / torch.sqrt(u ** (-2.0 / 3.0) - 1.0)
        d = torch.randn((nh_tot, 3), generator=gen, device=dev, dtype=torch.float32)
        d /= d.norm(dim=1, keepdim=True)
        pos[sel] = centres[hid] + r[:, None] * d
        vel[sel] = bulk[hid] + sig[hid][:, None] * torch.randn((nh_tot, 3), generator=gen, device=dev, dtype=torch.float32)
        del hid, sel, u, r, d
    pos -= torch.floor(pos / box) * box
    for j in range(3):
        col = pos[:, j]
        col[col >= float(box[j])] = 0.0          # fp32 rounding of values just below the period
    edges = None
    if slab is not None:
        rk, world = slab[0], slab[1]
        if len(slab) > 2 and slab[2] == "equal_count":
            # slab faces at the x quantiles: every slab holds n / world particles (up to ties at a face)
            xs = torch.sort(pos[:, 0]).values
            cut = [float(xs[(n * r) // world].item()) for r in range(1, world)]
            del xs
            edges = [0.0] + cut + [float(cx)]
        else:
            edges = [cx * r / world for r in range(world + 1)]
        keep = (pos[:, 0] >= edges[rk]) & (pos[:, 0] < edges[rk + 1])
        pos, vel = pos[keep], vel[keep]
    mass = torch.ones(pos.shape[0], device=dev, dtype=torch.float32)
    if return_edges:
        return pos.contiguous(), vel.contiguous(), mass, edges
    return pos.contiguous(), vel.contiguous(), mass
