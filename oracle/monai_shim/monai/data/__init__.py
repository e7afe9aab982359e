def decollate_batch(batch, detach=True, pad=True):
    return [batch[i] for i in range(batch.shape[0])]
