"""
Host -> device input feeder for the training loop (the reference does ``t.from_numpy(sample.image_data).unsqueeze(dim = 0).cuda()`` inside
the loop, pytorch/FasterRCNN/__main__.py:176-180: a synchronous copy from pageable memory in front of every step).

DeviceFeeder keeps two device slots per input and a copy stream: while step i runs, the inputs of step i + 1 are copied from page-locked
host memory by the copy engine, so the 7.7 MB per step (image + RPN ground-truth map at 600x1000) never sit on the compute stream's
critical path.  Usage:

  feeder = DeviceFeeder(device)
  feeder.submit(image_host, gt_map_host)              # inputs of the first step (pinned tensors)
  for sample in samples:
    image, gt_map = feeder.take()                     # device tensors of THIS step (the compute stream waits for their copy)
    feeder.submit(next_image_host, next_gt_map_host)  # starts the next step's copy behind this step's predecessors
    model.train_step(image_data = image, gt_rpn_map = gt_map, ...)
"""
import torch as t


class DeviceFeeder:
  def __init__(self, device):
    self.device = t.device(device)
    self.stream = t.cuda.Stream(device = self.device)
    self.slots = [None, None]            # per slot: list of device tensors
    self.ready = [None, None]            # per slot: event recorded on the copy stream after the slot's copies
    self.next_slot = 0                   # slot the next submit() fills
    self.pending = []                    # slots submitted and not yet taken, oldest first

  def submit(self, *host_tensors):
    """Queues the copy of one step's inputs (page-locked host tensors) into the free slot.  The copy is ordered behind everything the compute
    stream has been given so far -- which includes the last step that read this slot."""
    assert len(self.pending) < 2, "both slots are in flight: take() before submitting again"
    slot = self.next_slot
    main = t.cuda.current_stream(self.device)
    consumed = t.cuda.Event()
    consumed.record(main)
    if self.slots[slot] is None or len(self.slots[slot]) != len(host_tensors) or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(self.slots[slot], host_tensors)):
      self.slots[slot] = [t.empty(h.shape, dtype = h.dtype, device = self.device) for h in host_tensors]
    with t.cuda.stream(self.stream):
      self.stream.wait_event(consumed)
      for d, h in zip(self.slots[slot], host_tensors):
        d.copy_(h, non_blocking = True)
      done = t.cuda.Event()
      done.record(self.stream)
    self.ready[slot] = done
    self.pending.append(slot)
    self.next_slot = 1 - slot

  def take(self):
    """Device tensors of the oldest submitted step; the current (compute) stream waits for their copy."""
    assert self.pending, "nothing submitted"
    slot = self.pending.pop(0)
    t.cuda.current_stream(self.device).wait_event(self.ready[slot])
    return tuple(self.slots[slot])

  @staticmethod
  def bytes_per_step(*host_tensors):
    return sum(h.numel() * h.element_size() for h in host_tensors)
