"""Host-side helpers of the training-step path (mirror of the reference's utils/)."""
