"""The reference's engine surface (Engine.py:45-161; used by cldm_trt/ddim_hacked.py:140-169 and cldm_trt/cldm.py:368-384)
over this package's CUDA path: `Engine.infer(feed_dict, stream=None, use_cuda_graph=False) -> OrderedDict[name -> tensor]`
with the TensorRT engines' binding names and static shapes (onnx2trt_static_plugin.py:79-115):

  controlnet:  x_noisy [B,4,h,w], hint [B,3,8h,8w], timestep [B], context [B,77,768]  ->  13 control tensors
               (the dict lists the 4 inputs first, so `list(out.values())[4:17]` are the outputs, as the reference indexes)
  unet:        x_noisy, timestep, context, control0..control12                        ->  latent [B,4,h,w]
  decoder:     latent [B,4,h,w]                                                       ->  images [B,3,8h,8w]

There is no serialized plan: the "engine" is the live ControlLDM. Inputs are copied into pre-allocated binding tensors
(Engine.py:131-134), outputs live in tensors owned by the engine and are overwritten by the next call (callers `.clone()`
what they keep, like the reference). `stream` may be a torch.cuda.Stream, an object with a `.ptr` (polygraphy) or None."""
from collections import OrderedDict

import torch



class Engine:
    def __init__(self, model, kind, batch_size=1, latent_h=32, latent_w=48, text_maxlen=77):
        if kind not in ("controlnet", "unet", "decoder"):
            raise ValueError(f"unknown engine kind {kind!r}")
        self.model, self.kind = model, kind
        self.batch_size, self.latent_h, self.latent_w, self.text_maxlen = batch_size, latent_h, latent_w, text_maxlen
        self.tensors = OrderedDict()
        self.cuda_graph_instance = None
        self._static_out = None

    # ---- shapes (Engine.py:66-90, onnx2trt_static_plugin.py:79-115) ------------------------------------------------
    def control_shapes(self):
        """The 13 control tensor shapes for this latent size (export_onnx_all.py:242-256)."""
        unet = self.model.model.diffusion_model
        mc, b = unet.model_channels, self.batch_size
        shapes, h, w = [(b, mc, self.latent_h, self.latent_w)], self.latent_h, self.latent_w
        for level, mult in enumerate(unet.channel_mult):
            for _ in range(unet.num_res_blocks[level]):
                shapes.append((b, mc * mult, h, w))
            if level != len(unet.channel_mult) - 1:
                h, w = h // 2, w // 2
                shapes.append((b, mc * mult, h, w))
        shapes.append((b, mc * unet.channel_mult[-1], h, w))
        return shapes

    def shape_dict(self):
        b, h, w = self.batch_size, self.latent_h, self.latent_w
        unet = self.model.model.diffusion_model
        ctx_dim = self.model.control_model.input_blocks[1][1].transformer_blocks[0].attn2.to_k.in_features
        if self.kind == "controlnet":
            d = OrderedDict(x_noisy=(b, unet.in_channels, h, w), hint=(b, 3, 8 * h, 8 * w), timestep=(b,),
                            context=(b, self.text_maxlen, ctx_dim))
            for i, s in enumerate(self.control_shapes()):
                d[f"control{i}"] = s
            return d
        if self.kind == "unet":
            d = OrderedDict(x_noisy=(b, unet.in_channels, h, w), timestep=(b,), context=(b, self.text_maxlen, ctx_dim))
            for i, s in enumerate(self.control_shapes()):
                d[f"control{i}"] = s
            d["latent"] = (b, unet.out_channels, h, w)
            return d
        return OrderedDict(latent=(b, 4, h, w), images=(b, 3, 8 * h, 8 * w))

    def load(self):
        return self

    def activate(self, reuse_device_memory=None):
        return self

    def allocate_buffers(self, shape_dict=None, device=None):
        device = self.model.device if device is None else device
        for name, shape in self.shape_dict().items():
            if shape_dict and name in shape_dict:
                shape = shape_dict[name]
            dtype = torch.int64 if name == "timestep" else torch.float32
            self.tensors[name] = torch.empty(tuple(shape), dtype=dtype, device=device)
        return self

    # ---- execution -------------------------------------------------------------------------------------------------
    def _run(self):
        t, m = self.tensors, self.model
        if self.kind == "controlnet":
            outs = m.control_model(x=t["x_noisy"], hint=t["hint"], timesteps=t["timestep"], context=t["context"])
            for i, o in enumerate(outs):
                t[f"control{i}"].copy_(o)
        elif self.kind == "unet":
            control = [t[f"control{i}"] for i in range(13)]
            eps = m.model.diffusion_model(x=t["x_noisy"], timesteps=t["timestep"], context=t["context"], control=control,
                                          only_mid_control=m.only_mid_control)
            t["latent"].copy_(eps)
        else:
            t["images"].copy_(m.decode_first_stage(t["latent"]))

    @torch.no_grad()
    def infer(self, feed_dict, stream=None, use_cuda_graph=False):
        if not self.tensors:
            self.allocate_buffers()
        ts = self._torch_stream(stream)
        with torch.cuda.stream(ts):
            for name, buf in feed_dict.items():
                self.tensors[name].copy_(buf)       # (int32 timesteps of the ONNX export are widened by copy_)
            if use_cuda_graph:
                if self.cuda_graph_instance is None:
                    self._run()                      # inference before capture (Engine.py:145-149): packs / tunes
                    ts.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=ts):
                        self._run()
                    self.cuda_graph_instance = g
                else:
                    self.cuda_graph_instance.replay()
                ts.synchronize()                     # the reference synchronises after cudaGraphLaunch (Engine.py:142)
            else:
                self._run()
        return self.tensors

    def _torch_stream(self, stream):
        if stream is None:
            return torch.cuda.current_stream(self.model.device)
        if isinstance(stream, torch.cuda.Stream):
            return stream
        ptr = getattr(stream, "ptr", None)
        if ptr is not None:
            return torch.cuda.ExternalStream(int(ptr), device=self.model.device)
        raise TypeError(f"unsupported stream object {type(stream)}")
