"""Only the device-staging piece of the reference's data package is mirrored: the datasets,
degradation synthesis and samplers are the reference's CPU-side control plane and stay with it."""
