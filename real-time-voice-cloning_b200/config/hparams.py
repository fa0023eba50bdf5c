"""Constants of the path (reference: config/hparams.py:7-29 HParams, :38-51 sp, :220-285 wavernn_fatchord).
Only the fields the inference path reads are present."""
import ast


class HParams(object):
    """Attribute bag with `parse("k=v,k=v")` like the reference's (config/hparams.py:7-29)."""

    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __repr__(self):
        return "HParams(%s)" % ", ".join("%s=%r" % kv for kv in sorted(self.__dict__.items()))

    def parse(self, string):
        for item in filter(None, (s.strip() for s in string.split(","))):
            k, v = item.split("=", 1)
            self.__dict__[k.strip()] = ast.literal_eval(v.strip())
        return self


# signal processing constants shared with the synthesizer (config/hparams.py:38-51)
sp = HParams(sample_rate=16000, num_mels=80, hop_size=200, max_abs_value=4.0, preemphasis=0.97, preemphasize=True)

# fatchord WaveRNN (config/hparams.py:220-285); training-only fields omitted
wavernn_fatchord = HParams(
    mode="RAW", bits=10, mu_law=True, upsample_factors=(5, 5, 8),
    rnn_dims=512, fc_dims=512, compute_dims=128, res_out_dims=32 * 4, res_blocks=10, pad=2,
    use_sparsification=False, sparsity_target=0.90, sparse_group=4,
    gen_batched=True, gen_target=3000, gen_overlap=1500,
)

# RuntimeRacer's WaveRNN (config/hparams.py:355-421); training-only fields omitted
wavernn_runtimeracer = HParams(
    mode="RAW", bits=10, mu_law=True, upsample_factors=(5, 5, 8),
    rnn_dims=256, fc_dims=256, compute_dims=128, res_out_dims=64 * 2, res_blocks=10, pad=2,
    use_sparsification=False, sparsity_target=0.90, sparse_group=4,
    gen_batched=True, gen_target=6000, gen_overlap=1000,
)

# geneing's WaveRNN (config/hparams.py:288-352); training-only fields omitted
wavernn_geneing = HParams(
    mode="BITS", bits=10, mu_law=False, upsample_factors=(4, 5, 10),
    rnn_dims=256, fc_dims=128, compute_dims=64, res_out_dims=32 * 2, res_blocks=3, pad=2,
    use_sparsification=False, sparsity_target=0.90, sparse_group=4,
    gen_batched=True, gen_target=3000, gen_overlap=1500,
)
