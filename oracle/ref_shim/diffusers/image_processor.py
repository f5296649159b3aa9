class VaeImageProcessor:  # only constructed by the pipeline's __init__; not on the denoise path
    def __init__(self, *a, **k):
        self.args = (a, k)

    def preprocess(self, *a, **k):
        raise NotImplementedError("VAE image preprocessing is out of scope for the oracle harness")
