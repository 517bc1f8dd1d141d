"""ORACLE (test infrastructure only - never imported by the product path).

Line-by-line numpy restatement of the per-frame volume pipelines of the reference:
  out_generate3d   GAN/multipassGAN-out.py:390-618 (1-3 chained generators, transposeAxis 0-3)
  apply_4x_pass    GAN/multipassGAN-4x.py:1090-1169 (one generator pass, upsampling modes 1/2/3)
  two_pass_4x      GAN/example_run_output.py:6,8 (pass 1 mode 2 -> .uni -> pass 2 mode 1)
The networks are injected as callables on flat rows exactly like `sess.run(sampler, feed_dict)`, so
the axis / channel bookkeeping can be tested with identity "networks".  PINNED bit-exactly against the
reference's own generate3DUniForNewNetwork executed with stand-in networks (tests/golden/pipeline.npz,
tests/test_golden.py).
"""
import numpy as np
import scipy.ndimage


def _zoom(a, z):
    return scipy.ndimage.zoom(a, z, order=1, mode="constant", cval=0.0)


def _swap(b, i, j):
    """The reference's three-line np.copy channel swap (e.g. GAN/multipassGAN-out.py:473-475)."""
    t = np.copy(b[:, :, :, i:i + 1])
    b[:, :, :, i:i + 1] = np.copy(b[:, :, :, j:j + 1])
    b[:, :, :, j:j + 1] = t


def _add_adj(b, C):
    """GAN/multipassGAN-out.py:423-436: channels C, C+1 = density of slice i-1 / i+1 (zeros at the ends)."""
    b = np.concatenate((b, np.zeros_like(b[:, :, :, 0:1])), axis=3)
    b = np.concatenate((b, np.zeros_like(b[:, :, :, 0:1])), axis=3)
    n = b.shape[0]
    for i in range(n):
        if i == 0:
            b[i, :, :, C] = 0
            b[i, :, :, C + 1] = b[i + 1, :, :, 0]
        elif i == n - 1:
            b[i, :, :, C] = b[i - 1, :, :, 0]
            b[i, :, :, C + 1] = 0
        else:
            b[i, :, :, C] = b[i - 1, :, :, 0]
            b[i, :, :, C + 1] = b[i + 1, :, :, 0]
    return b


def _run_batches(net, xs, n_input, batch, ys=None, n_output=None):
    """`for j in range(N//B): sess.run(...)`; the trailing N mod B slices are dropped (App. D.4)."""
    rows = []
    for j in range(0, xs.shape[0] // batch):
        xb = xs[j * batch:(j + 1) * batch].reshape(-1, n_input)
        if ys is None:
            res = net(xb)
        else:
            res = net(xb, ys[j * batch:(j + 1) * batch].reshape(-1, n_output))
        rows.extend(np.asarray(res))
    return rows


def out_pass1_input(x, L, S, u, C, transposeAxis, add_adj):
    """GAN/multipassGAN-out.py:398-436. x: [L,L,L,C] (z,y,x,c)."""
    if transposeAxis == 1:
        b = np.reshape(_zoom(x, [1, u, 1, 1]), [-1, S, L, C])
        b = np.reshape(b.transpose(1, 0, 2, 3), (-1, L, L, C))
        _swap(b, 3, 2)
    elif transposeAxis == 2:
        b = np.reshape(_zoom(x, [1, 1, u, 1]), [-1, L, S, C])
        b = np.reshape(b.transpose(2, 1, 0, 3), (-1, L, L, C))
        _swap(b, 3, 1)
    elif transposeAxis == 3:
        b = np.reshape(_zoom(x, [1, 1, u, 1]), [-1, L, S, C])
        b = np.reshape(b.transpose(2, 0, 1, 3), (-1, L, L, C))
        t = np.copy(b[:, :, :, 3:4])
        t2 = np.copy(b[:, :, :, 2:3])
        b[:, :, :, 3:4] = np.copy(b[:, :, :, 1:2])
        b[:, :, :, 2:3] = t
        b[:, :, :, 1:2] = t2
    else:
        b = np.reshape(_zoom(x, [u, 1, 1, 1]), [-1, L, L, C])
    if add_adj:
        b = _add_adj(b, C)
    return b


def out_pass2_input(x, L, S, u, C, transposeAxis):
    """GAN/multipassGAN-out.py:464-485 (add_adj_idcs2 is broken in the reference, App. D.7)."""
    if transposeAxis == 3:
        b = np.reshape(_zoom(x, [1, u, 1, 1]), [-1, S, L, C])
        b = np.reshape(b.transpose(1, 0, 2, 3), (-1, L, L, C))
        _swap(b, 3, 2)
    elif transposeAxis == 0:
        b = np.reshape(_zoom(x, [1, 1, u, 1]), [-1, L, S, C])
        b = np.reshape(b.transpose(2, 1, 0, 3), (-1, L, L, C))
        _swap(b, 3, 1)
    elif transposeAxis == 1:
        b = np.reshape(_zoom(x, [1, 1, u, 1]), [-1, L, S, C])
        b = np.reshape(b.transpose(2, 0, 1, 3), (-1, L, L, C))
        t = np.copy(b[:, :, :, 3:4])
        t2 = np.copy(b[:, :, :, 2:3])
        b[:, :, :, 3:4] = np.copy(b[:, :, :, 1:2])
        b[:, :, :, 2:3] = t
        b[:, :, :, 1:2] = t2
    else:
        b = np.reshape(_zoom(x, [u, 1, 1, 1]), [-1, L, L, C])
    return b


def out_pass3_input(x, L, S, u, C, transposeAxis):
    """GAN/multipassGAN-out.py:526-547 (transposeAxis 2 indexes channel 13 in the reference: dead branch)."""
    if transposeAxis == 0:
        b = np.reshape(_zoom(x, [1, u, 1, 1]), [-1, S, L, C])
        b = np.reshape(b.transpose(1, 0, 2, 3), (-1, L, L, C))
        _swap(b, 3, 2)
    elif transposeAxis == 3:
        b = np.reshape(_zoom(x, [1, 1, u, 1]), [-1, L, S, C])
        b = np.reshape(b.transpose(0, 2, 1, 3), (-1, L, L, C))
        _swap(b, 2, 1)
    elif transposeAxis == 2:
        raise NotImplementedError("reference indexes channel 13 here (App. D.7)")
    else:
        b = np.reshape(_zoom(x, [u, 1, 1, 1]), [-1, L, L, C])
    return b


def out_generate3d(x, u, net1=None, net2=None, net3=None, transposeAxis=0, add_adj_idcs1=False,
                   threshold=True, return_intermediate=False):
    """GAN/multipassGAN-out.py:390-618 for one frame. x: [L,L,L,C] float32 (velocities already
    scaled, :138). net*: callables (x_rows[, y_rows]) -> rows [B, S*S]. Returns [S,S,S] (z,y,x)."""
    x = np.asarray(x, dtype=np.float32)
    L = x.shape[0]
    C = x.shape[3]
    S = L * u
    n_input = L * L * C
    n_output = S * S
    inter = {}
    dim_output = None
    if net1 is not None:
        b = out_pass1_input(x, L, S, u, C, transposeAxis, add_adj_idcs1)
        rows = _run_batches(net1, b, b.shape[1] * b.shape[2] * b.shape[3], 8)
        dim_output = np.copy(np.array(rows).reshape(S, S, S)).transpose(2, 1, 0)
        inter["pass1"] = np.ascontiguousarray(dim_output)
    if net2 is not None:
        b = out_pass2_input(x, L, S, u, C, transposeAxis)
        rows = _run_batches(net2, b, n_input, 2, dim_output, n_output)
        dim_output = np.array(rows).reshape(S, S, S).transpose(1, 2, 0)
        inter["pass2"] = np.ascontiguousarray(dim_output)
    if net3 is not None:
        b = out_pass3_input(x, L, S, u, C, transposeAxis)
        rows = _run_batches(net3, b, n_input, 2, dim_output, n_output)
        dim_output = np.array(rows).reshape(S, S, S)
        inter["pass3"] = np.ascontiguousarray(dim_output)
    # :587-590 -- keyed on load_model_no_2 / _1, i.e. on whether nets 2 / 1 are configured
    if net2 is not None:
        dim_output = dim_output.transpose(2, 0, 1)
    if net1 is not None:
        dim_output = dim_output.transpose(2, 1, 0)
    dim_output = np.array(dim_output, dtype=np.float32)
    if threshold:  # :612-615 (genUni 1)
        dim_output[dim_output < 0.0005] = 0
    if return_intermediate:
        return dim_output, inter
    return dim_output


def apply_4x_pass(net, u, mode, x_3d, x_2=None, threshold=True):
    """GAN/multipassGAN-4x.py:1090-1169 for one frame (upsampleFirst=1).
    mode 2: x_3d = [L,L,L,C] low-res (d,vx,vy,vz).
    mode 1/3: x_3d = [L,L,L,3] velocities already multiplied by upRes (:277-278), x_2 = [S,S,S,1]
    density written by the previous pass.  Returns [S,S,S] float32 (z,y,x)."""
    x_3d = np.asarray(x_3d, dtype=np.float32)
    L = x_3d.shape[0]
    S = L * u
    if mode in (1, 3):
        tile = _zoom(x_3d, [u, u, u, 1])
    else:
        tile = x_3d
    if mode == 2:
        C = x_3d.shape[3]
        b = np.reshape(_zoom(tile, [u, 1, 1, 1]), [-1, L, L, C])
        n_input = L * L * C
    elif mode == 1:
        C = 4
        b = np.reshape(np.concatenate((x_2, tile), axis=3), [-1, S, S, S, C]).transpose((0, 3, 1, 2, 4)).reshape(
            [-1, S, S, C])
        b = np.array(b)
        _swap(b, 2, 3)
        _swap(b, 3, 1)
        n_input = S * S * C
    elif mode == 3:
        C = 4
        b = np.reshape(np.concatenate((x_2, tile), axis=3), [-1, S, S, S, C]).transpose((0, 2, 1, 3, 4)).reshape(
            [-1, S, S, C])
        b = np.array(b)
        _swap(b, 2, 3)
        n_input = S * S * C
    else:
        raise NotImplementedError("upsampling_mode 0 is not on the benchmarked path")
    rows = _run_batches(net, b, n_input, 8)
    out = np.array(rows).reshape(S, S, S)
    if mode == 1:
        out = out.transpose(1, 2, 0)
    elif mode == 3:
        out = out.transpose(1, 0, 2)
    out = np.array(out, dtype=np.float32)
    if threshold:  # :1155-1157 (genUni 1)
        out[out < 0.0005] = 0
    return out


def two_pass_4x(net_pass1, net_pass2, x, u=4, velScale=1.0, return_intermediate=False):
    """The shipped 4x multi-pass recipe (GAN/example_run_output.py:6,8): pass 1 `upsamplingMode 2
    upsampledData 0`, its thresholded .uni output is the `x_2` of pass 2 `upsamplingMode 1
    upsampledData 1`. x: [L,L,L,4]."""
    x = np.array(x, dtype=np.float32)
    x1 = np.copy(x)
    x1[..., 1:4] = velScale * x1[..., 1:4]  # GAN/multipassGAN-4x.py:283 (upsampled_data 0 branch)
    p1 = apply_4x_pass(net_pass1, u, 2, x1)
    vel = x[..., 1:4] * u  # :277-278
    vel[..., 1:4] = velScale * vel[..., 1:4]  # :283 -- App. D.10: hits only vy,vz of the 3-channel array
    p2 = apply_4x_pass(net_pass2, u, 1, vel, x_2=p1[..., None])
    if return_intermediate:
        return p2, p1
    return p2
