"""Mirror of models/gan/stylegan2/ (rosinality-style StyleGAN2 used by train_stylegan2*.py) on the sm_100a kernels."""
