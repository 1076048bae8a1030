"""What every loss offers to the agents (src/losses/abstract_loss.py:20-41): it is callable, and the two hooks the
optimisation loop calls before / after an episode step default to no-ops."""
from ..core import module


class abstract_loss(module):
    def __call__(self, *args, **kwargs):
        return self.evaluate(*args, **kwargs)

    def evaluate(self, *args, **kwargs):
        raise NotImplementedError("{} does not implement an evaluate() method".format(self.__class__))

    def prepare(self, sess, inputs_old, outputs_old, parallelism, **kwargs):
        """feed construction hook; subclasses draw their base samples here"""

    def update(self, sess, inputs_old, outputs_old, **kwargs):
        """state refresh hook after new observations"""
