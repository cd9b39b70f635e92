"""Oracle (test infrastructure): energy/force prediction.

NumPy restatement of sgdml/predict.py:84-245 (_predict_wkr), the permuted caches of
predict.py:424-441, set_alphas predict.py:551-601 and the output scaling of
predict.py:1286-1288, including the energy-constraint terms (predict.py:219-229) and lattices (predict.py:332-334).
"""

import multiprocessing as mp

import numpy as np

from . import desc as odesc


class Predictor(object):
    """Host-only stand-in for GDMLPredict(model, use_torch=False)."""

    def __init__(self, model):
        assert model['type'] in ('m', b'm')  # predict.py:326-328
        self.n_atoms = int(np.asarray(model['z']).shape[0])
        self.sig = float(model['sig'])  # NumPy path uses sig as stored (predict.py:346)
        self.std = float(model['std']) if 'std' in model else 1.0
        self.c = float(model['c'])
        self.n_perms = int(np.asarray(model['perms']).shape[0])
        self.tril_perms_lin = np.asarray(model['tril_perms_lin'])
        self.n_train = model['R_desc'].shape[1]
        self.R_d_desc = None
        self.R_desc_train = None
        self.lat_and_inv = None
        if 'lattice' in model:  # predict.py:332-334
            lat = np.asarray(model['lattice'], dtype=np.float64)
            self.lat_and_inv = (lat, np.linalg.inv(lat))
        # energy constraints in the kernel: one coefficient per training point, repeated over the permutations
        # (predict.py:443-447)
        self.alphas_E_lin = None
        if 'alphas_E' in model:
            self.alphas_E_lin = np.tile(np.asarray(model['alphas_E'], dtype=np.float64)[:, None], (1, self.n_perms)).ravel()

        # predict.py:426-441: caches with row k = m*S + p
        self.R_desc_perms = self._perm_cache(np.asarray(model['R_desc']).T)
        self.R_d_desc_alpha_perms = self._perm_cache(np.asarray(model['R_d_desc_alpha']))

    def _perm_cache(self, X):
        """(M, D) -> (M*S, D): np.tile(X, S)[:, tril_perms_lin] reshaped Fortran-wise."""
        M = X.shape[0]
        S = self.n_perms
        return (
            np.tile(X, S)[:, self.tril_perms_lin]
            .reshape(M, S, -1, order='F')
            .reshape(M * S, -1)
        )

    def set_R_desc(self, R_desc):  # predict.py:511-525
        self.R_desc_train = R_desc

    def set_R_d_desc(self, R_d_desc):  # predict.py:527-549
        self.R_d_desc = R_d_desc

    def set_alphas(self, alphas_F):
        """predict.py:551-601 (force coefficients only)."""
        assert self.R_d_desc is not None
        R_d_desc_alpha = odesc.d_desc_dot_vec(self.R_d_desc, alphas_F.reshape(-1, 3 * self.n_atoms))
        self.R_d_desc_alpha_perms = self._perm_cache(R_d_desc_alpha)

    def _raw(self, r_desc, r_d_desc):
        """One geometry, unscaled [E, F(3N)] (predict.py:150-245, single chunk)."""
        sig = self.sig
        sig_inv = 1.0 / sig
        mat52_base_fact = 5.0 / (3 * sig**3)
        diag_scale_fact = 5.0 / sig
        sqrt5 = np.sqrt(5.0)

        diff = r_desc[None, :] - self.R_desc_perms  # predict.py:199-203
        norm = sqrt5 * np.sqrt(np.sum(diff * diff, axis=1))  # predict.py:204
        base = np.exp(-norm * sig_inv) * mat52_base_fact  # predict.py:206-207
        a_x2 = np.einsum('ji,ji->j', diff, self.R_d_desc_alpha_perms)  # predict.py:208-210

        Fd = (a_x2 * base).dot(diff) * diag_scale_fact  # predict.py:212
        base = base * (norm + sig)  # predict.py:213
        Fd -= base.dot(self.R_d_desc_alpha_perms)  # predict.py:214
        E = a_x2.dot(base)  # predict.py:217
        if self.alphas_E_lin is not None:  # predict.py:219-229
            Fd += self.alphas_E_lin.dot(diff * base[:, None])
            K_ee = (1 + (norm * sig_inv) * (1 + norm / (3 * sig))) * np.exp(-norm * sig_inv)
            E += K_ee.dot(self.alphas_E_lin)

        out = np.empty(3 * self.n_atoms + 1)
        out[0] = E
        out[1:] = odesc.vec_dot_d_desc(r_d_desc, Fd)[0]  # predict.py:240-243
        return out

    def predict(self, R=None, return_E=True):
        """predict.py:1146-1294, CPU branch: R (B, 3N) -> E (B,), F (B, 3N).
        R=None evaluates on the cached training descriptors (predict.py:1219-1235)."""
        if R is None:
            assert self.R_desc_train is not None and self.R_d_desc is not None
            R_desc, R_d_desc = self.R_desc_train, self.R_d_desc
        else:
            R = np.asarray(R, dtype=np.float64)
            if R.ndim == 1:
                R = R[None, :]  # predict.py:1183-1184
            R_desc, R_d_desc = odesc.from_R(R.reshape(R.shape[0], -1), self.lat_and_inv)
        E_F = np.array([self._raw(x, g) for x, g in zip(R_desc, R_d_desc)])
        E_F = E_F.reshape(-1, 3 * self.n_atoms + 1) * self.std  # predict.py:1286
        F = E_F[:, 1:]
        E = E_F[:, 0] + self.c  # predict.py:1288
        return (E, F) if return_E else (F,)


_WKR_PRED = None


def _mp_init(model):
    global _WKR_PRED
    try:  # one BLAS thread per worker process: the pool itself covers the host threads
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=1)
    except Exception:
        pass
    _WKR_PRED = Predictor(model)


def _mp_predict(R_chunk):
    return _WKR_PRED.predict(R_chunk)


class ParallelPredictor(object):
    """Bulk prediction over all host threads -- the reference's ``bulk_mp`` mode
    (predict.py:1237-1257): a pool of worker processes, whole geometries per task.  Workers are forked
    from a clean ``forkserver`` process that has NumPy and this module preloaded: safe after CUDA
    initialisation in the parent (no CUDA state is inherited) and much cheaper than 128 ``spawn``
    interpreters each importing NumPy."""

    def __init__(self, model, n_procs):
        self.n_procs = max(int(n_procs), 1)
        self.pool = None
        if self.n_procs > 1:
            ctx = mp.get_context('forkserver')
            try:
                ctx.set_forkserver_preload(['numpy', 'oracle.predict'])
            except Exception:
                pass
            self.pool = ctx.Pool(self.n_procs, initializer=_mp_init, initargs=(model,))
        else:
            self.single = Predictor(model)

    def predict(self, R):
        R = np.asarray(R, dtype=np.float64)
        if self.pool is None:
            return self.single.predict(R)
        n_chunks = max(1, min(len(R), self.n_procs * 4))
        res = self.pool.map(_mp_predict, np.array_split(R, n_chunks), chunksize=1)
        return np.concatenate([r[0] for r in res]), np.concatenate([r[1] for r in res])

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def predict_parallel(model, R, n_procs):
    with ParallelPredictor(model, n_procs) as pp:
        return pp.predict(R)
