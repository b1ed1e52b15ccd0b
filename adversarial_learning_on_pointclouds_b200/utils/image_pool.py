"""History buffer of generated maps with the reference's interface
(utils/image_pool.py:10-55, the CycleGAN image pool)."""
import random

import torch


class ImagePool:
    def __init__(self, pool_size):
        self.pool_size = pool_size
        self.num_imgs = 0
        self.images = []

    def query(self, images):
        """pool_size 0: pass-through.  Otherwise, per sample: fill the buffer
        first; once full, with probability 1/2 swap the sample with a random
        stored one and return the stored one."""
        if self.pool_size == 0:
            return images
        out = []
        for image in images:
            image = torch.unsqueeze(image.data, 0)
            if self.num_imgs < self.pool_size:
                self.num_imgs += 1
                self.images.append(image)
                out.append(image)
            elif random.uniform(0, 1) > 0.5:
                j = random.randint(0, self.pool_size - 1)
                out.append(self.images[j].clone())
                self.images[j] = image
            else:
                out.append(image)
        return torch.cat(out, 0).requires_grad_(True)
