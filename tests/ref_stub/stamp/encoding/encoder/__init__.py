from abc import ABC, abstractmethod


class Encoder(ABC):
    def __init__(self, model, identifier, precision, required_extractors):
        self.model = model
        self.identifier = identifier
        self.precision = precision
        self.required_extractors = required_extractors

    def encode_slides_(self, **kwargs):
        return ("walked", self.identifier, kwargs)

    @abstractmethod
    def _generate_slide_embedding(self, feats, device, **kwargs): ...

    @abstractmethod
    def _generate_patient_embedding(self, feats_list, device, **kwargs): ...
