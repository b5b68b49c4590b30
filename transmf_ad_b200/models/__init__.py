"""Drop-in mirror of the reference ``models`` package (same module / class names, constructor and forward
signatures, ``state_dict`` keys) running on the B200 kernels.  Put the directory that contains this package
first on ``sys.path`` and the reference training scripts import it unchanged:

    from models.mymodel import model_ad, model_CNN_ad, model_single
"""
