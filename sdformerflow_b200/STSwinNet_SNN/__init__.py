"""Host-side mirror of the reference's ``models/STSwinNet_SNN`` package: same class names,
constructor arguments, config keys and ``state_dict`` layout, with the
hot path running on the sm_100a kernels of libsdf_b200."""
