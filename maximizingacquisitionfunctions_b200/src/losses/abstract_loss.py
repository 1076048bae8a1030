from ..core import module


class abstract_loss(module):
    """Interface of the losses (reference: src/losses/abstract_loss.py:20-41)."""

    def evaluate(self, *args, **kwargs):
        raise NotImplementedError("{} does not implement an evaluate() method".format(self.__class__))

    def __call__(self, *args, **kwargs):
        return self.evaluate(*args, **kwargs)

    def prepare(self, sess, inputs_old, outputs_old, parallelism, **kwargs):
        pass

    def update(self, sess, inputs_old, outputs_old, **kwargs):
        pass
