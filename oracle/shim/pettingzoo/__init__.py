"""oracle shim: pettingzoo.AECEnv.last() as documented [UPSTREAM-UNVERIFIED]."""


class AECEnv:
    def __init__(self, *a, **k):
        pass

    def last(self, observe=True):
        a = self.agent_selection
        obs = self.observe(a) if observe else None
        return obs, self._cumulative_rewards[a], self.terminations[a], self.truncations[a], self.infos[a]

    def close(self):
        pass
