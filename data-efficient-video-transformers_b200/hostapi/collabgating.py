"""Collaborative gating fusion (src/models/collabgating.py:3-56).

The reference walks nested Python lists - per clip, per scene, per expert - and issues O(E^2) ``Linear(2048, 2048)``
calls on ``[1, 2048]`` vectors.  Here every (clip, scene) is one row and the E(E+1)/2 + E projections collapse into
THREE tensor-core GEMMs on stacked rows:
    C  = P(pad(X))     for all experts         [E*N, 2048]
    PC = P(C_0..E-2)   the re-projection the reference applies to experts it has already projected and pushed back
                       on its work list (collabgating.py:49)                                   [(E-1)*N, 2048]
    A  = P(T)          with T_i = (E-1) C_i + sum_{j>i} C_j + sum_{j<i} PC_j                    [E*N, 2048]
then out = sum_i C_i * sigmoid(C_i + A_i)  (ContextGating's GLU, :83-85) and normalize(geu.fc(out)).
Everything between the GEMMs is one fused kernel each, forward and backward (csrc/collab.cu): the nearest-neighbour stretch +
cast of the inputs (``tvt_stretch_cast``), T (``tvt_collab_mix``), the gate sum (``tvt_collab_gate``) and the L2
normalisation (``tvt_l2norm``) - no eager torch arithmetic is left on the path.
Parameter names / shapes are the reference's (``projection.*``, ``geu.fc.*``)."""
import torch
import torch.nn as nn

from .. import ops
from ..functions import CollabGateFn, CollabMixFn, L2NormFn, LinearFn


class GatedEmbeddingUnit(nn.Module):
    def __init__(self, input_dimension, output_dimension, use_bn=False):
        super().__init__()
        self.fc = nn.Linear(input_dimension, output_dimension)


class CollaborativeGating(nn.Module):
    def __init__(self, precision="bf16"):
        super().__init__()
        self.proj_input = 2048
        self.proj_embedding_size = 2048
        self.projection = nn.Linear(self.proj_input, self.proj_embedding_size)
        self.geu = GatedEmbeddingUnit(self.proj_input, 1024, False)
        self.mode = ops.Mode(precision)

    def pad(self, x):
        """Nearest-neighbour stretch of a narrower expert to 2048 (collabgating.py:12-16): out[i] = in[floor(i * D / 2048)]."""
        D = x.shape[-1]
        if D == self.proj_input:
            return x
        idx = (torch.arange(self.proj_input, device=x.device) * D) // self.proj_input
        return x.index_select(-1, idx)

    @staticmethod
    def from_nested(batch):
        """The reference's input layout (list of clips of scenes of experts of [1, D_e]) -> list over experts of [B, S, D_e]."""
        E = len(batch[0][0])
        return [torch.stack([torch.stack([scene[e].reshape(-1) for scene in clip]) for clip in batch]) for e in range(E)]

    def forward(self, xs):
        if isinstance(xs[0], (list, tuple)):
            xs = self.from_nested(xs)
        E = len(xs)
        if E < 2 or E > 8:
            raise ValueError("CollaborativeGating needs between two and eight experts")
        B, S = xs[0].shape[:2]
        N = B * S
        m = self.mode
        D = self.proj_input
        if not xs[0].is_cuda:
            raise ops.TvtError("input tensor is not on a CUDA device: this path has no CPU implementation")
        P = lambda rows: LinearFn.apply(m, rows, self.projection.weight, self.projection.bias)
        # every expert stretched to 2048 (nearest neighbour, collabgating.py:12-16) and cast, straight into its slice of the
        # stacked operand: no index_select / cat / cast passes
        X = torch.empty(E * N, D, dtype=m.dtype, device=xs[0].device)
        for e, x in enumerate(xs):
            ops.stretch_cast(x.reshape(N, -1), X[e * N:(e + 1) * N])
        C = P(X)                                                           # [E*N, 2048]
        PC = P(C[: (E - 1) * N])                                           # experts 0..E-2, projected twice (:49)
        T = CollabMixFn.apply(C.view(E, N, D), PC.view(E - 1, N, D))       # pairwise-sum bookkeeping, one pass
        A = P(T.view(E * N, D))
        gated = CollabGateFn.apply(C.view(E, N, D), A.view(E, N, D))       # sum_i GLU(cat(C_i, C_i + A_i))
        out = LinearFn.apply(m, gated, self.geu.fc.weight, self.geu.fc.bias)
        return L2NormFn.apply(out).view(B, S, -1)
