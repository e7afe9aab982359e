from abc import ABC, abstractmethod


class Inferer(ABC):
    @abstractmethod
    def __call__(self, inputs, network, *args, **kwargs):
        raise NotImplementedError


class SimpleInferer(Inferer):
    def __call__(self, inputs, network, *args, **kwargs):
        return network(inputs, *args, **kwargs)
