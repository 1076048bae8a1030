from .search_agent import search_agent


class random_search(search_agent):
    """Uniform random suggestions (reference: src/agents/random_search.py:19-23)."""

    def suggest_inputs(self, num_suggest, inputs_old, outputs_old=None, *args, **kwargs):
        input_dim = inputs_old.shape[-1]
        return self.sample_inputs([num_suggest, input_dim]), None
