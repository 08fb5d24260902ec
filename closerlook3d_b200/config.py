"""Configuration keys the local-aggregation constructors read.

The reference passes a global EasyDict (`/root/reference/pytorch/utils/config.py:4-103`) into every
`LocalAggregation(in_channels, out_channels, radius, nsample, config)`.  The drop-in modules in this
package accept that same object unchanged (attribute access is all they need).  `la_config()` builds a
stand-alone object with the same keys and the same defaults (config.py:27,33,77-103) for use without the
reference tree (tests, bench).
"""


class AttrDict(dict):
    """dict with attribute access (the subset of EasyDict behaviour the constructors rely on)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def la_config(local_aggregation_type="pospool", **overrides):
    """Defaults as in the reference's utils/config.py; nested dict overrides, e.g.
    la_config('pospool', pospool={'position_embedding': 'sin_cos', 'reduction': 'avg'})."""
    c = AttrDict()
    c.bn_momentum = 0.1
    c.density_parameter = 5.0
    c.local_aggregation_type = local_aggregation_type
    c.pospool = AttrDict(position_embedding="xyz", reduction="sum", output_conv=False)
    c.adaptive_weight = AttrDict(weight_type="dp", num_mlps=1, shared_channels=1, weight_softmax=False,
                                 reduction="avg", output_conv=False)
    c.pointwisemlp = AttrDict(feature_type="dp_fj", num_mlps=1, reduction="max")
    c.pseudo_grid = AttrDict(fixed_kernel_points="center", KP_influence="linear", KP_extent=1.0,
                             num_kernel_points=15, convolution_mode="sum", output_conv=False)
    for k, v in overrides.items():
        if isinstance(v, dict):
            c[k].update(v)
        else:
            c[k] = v
    return c


# The five BASELINE.json configurations: family settings as in the shipped cfgs/*.yaml.
def baseline_config(i):
    """i in 1..5 -> dict(name, family cfg, B, N, K, C, gpus)."""
    table = {
        1: dict(name="c1 ModelNet40 PosPool xyz avg", la="pospool",
                over=dict(pospool=dict(position_embedding="xyz", reduction="avg")), B=2, N=1024, K=16, C=66, gpus=1),
        2: dict(name="c2 ModelNet40 PointWiseMLP dp_fi_df max", la="pointwisemlp",
                over=dict(pointwisemlp=dict(feature_type="dp_fi_df", num_mlps=1, reduction="max")),
                B=32, N=1024, K=32, C=72, gpus=1),
        3: dict(name="c3 S3DIS PseudoGrid linear 15kp sum", la="pseudo_grid",
                over=dict(), B=8, N=15000, K=26, C=72, gpus=1),
        4: dict(name="c4 PartNet AdaptiveWeight dp fc1 avg", la="adaptive_weight",
                over=dict(adaptive_weight=dict(weight_type="dp", num_mlps=1, shared_channels=1, reduction="avg")),
                B=32, N=10000, K=32, C=72, gpus=8),
        5: dict(name="c5 S3DIS PosPool sin_cos avg width x2", la="pospool",
                over=dict(pospool=dict(position_embedding="sin_cos", reduction="avg")),
                B=64, N=40000, K=40, C=144, gpus=8),
    }
    t = dict(table[i])
    t["cfg"] = la_config(t["la"], **t["over"])
    t["index"] = i
    return t
