from .layerstack import LayerStack
