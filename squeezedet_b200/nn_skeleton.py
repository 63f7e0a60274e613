"""Neural-network model base class — the drop-in for reference ``src/nn_skeleton.py``.

Same constructor (`ModelSkeleton(mc)`), same layer-constructor methods with the
same argument meaning (`_conv_layer`, `_conv_bn_layer`, `_pooling_layer`), same
graph attributes (`image_input`, `preds`, `det_boxes`, `det_probs`, `det_class`,
`model_params`, `model_size_counter`, `flop_counter`, `activation_counter`) and the
same `filter_prediction(boxes, probs, cls_idx)`.  Instead of building a TF-1.0
graph, each constructor records the layer into a `libsqdet_b200` engine plan
(C ABI, `include/sqdet_b200.h`); the arithmetic runs in hand-written sm_100a
kernels.  There is no TensorFlow and no CPU fallback.

What replaces `sess.run([model.det_boxes, model.det_probs, model.det_class],
feed_dict={model.image_input: images})` (reference src/demo.py:193-195):
  * `Session().run(fetches, feed_dict)`           — same call shape, or
  * `model.detect(images)`                         — direct,
  * `model.detect_filtered(images)`                — detect + filter_prediction on
    every image in one GPU pass (what demo.py:193-199 / eval.py:75-87 do in two
    steps with a Python loop).
Training-only graph pieces (loss / train / viz graphs, `_fc_layer`, the input
FIFO queue; nn_skeleton.py:285-372,589-694) are out of scope of this engine.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import SqdetError  # noqa: F401


class GraphTensor:
  """Opaque handle to a tensor of the engine plan (stands in for a tf.Tensor)."""

  def __init__(self, model, name, tid=None, shape=None, fetch=None):
    self.model = model
    self.name = name
    self.id = tid
    self.shape = shape            # (B, H, W, C) for activations
    self.fetch = fetch            # 'det_boxes' | 'det_probs' | 'det_class' | None

  def get_shape(self):
    return self.shape

  def __repr__(self):
    return 'GraphTensor(%r, shape=%r)' % (self.name, self.shape)


class GraphParam:
  """Handle to one model parameter (stands in for a tf.Variable in
  `model.model_params`, which the reference hands to tf.train.Saver)."""

  def __init__(self, model, name, shape):
    self.model = model
    self.name = name
    self.shape = tuple(shape)

  def __repr__(self):
    return 'GraphParam(%r, shape=%r)' % (self.name, self.shape)


class Session:
  """Minimal stand-in for tf.Session so reference-style callers keep reading
  `sess.run(fetches, feed_dict={model.image_input: images})`."""

  def __init__(self, *args, **kwargs):
    pass

  def __enter__(self):
    return self

  def __exit__(self, *exc):
    return False

  def run(self, fetches, feed_dict=None):
    single = isinstance(fetches, GraphTensor)
    flist = [fetches] if single else list(fetches)
    if not flist:
      return []
    model = flist[0].model
    out = model.run(flist, feed_dict or {})
    return out[0] if single else out


class ModelSkeleton:
  """Base class of NN detection models (reference nn_skeleton.py:72)."""

  def __init__(self, mc, gpu_id=0, math_mode=None):
    self.mc = mc
    self.gpu_id = int(gpu_id)
    # dropout keep probability (nn_skeleton.py:78); inference => identity
    self.keep_prob = 0.5 if mc.IS_TRAINING else 1.0
    if mc.IS_TRAINING:
      raise NotImplementedError(
          'squeezedet_b200 is an inference engine: mc.IS_TRAINING must be False')
    if math_mode is None:
      math_mode = getattr(mc, 'MATH_MODE', _lib.MATH_TF32X3_TC)
    self.math_mode = int(math_mode)
    cfg = _lib.Config(
        batch_size=int(mc.BATCH_SIZE), image_height=int(mc.IMAGE_HEIGHT),
        image_width=int(mc.IMAGE_WIDTH), classes=int(mc.CLASSES),
        anchors_per_grid=int(mc.ANCHOR_PER_GRID),
        top_n_detection=int(mc.TOP_N_DETECTION),
        prob_thresh=float(mc.PROB_THRESH), nms_thresh=float(mc.NMS_THRESH),
        exp_thresh=float(mc.EXP_THRESH),
        batch_norm_epsilon=float(mc.BATCH_NORM_EPSILON),
        math_mode=self.math_mode, max_dets=int(getattr(mc, 'MAX_DETS', 0)))
    self._lib = _lib.load()
    handle = C.c_void_p()
    _lib.check(self._lib.sqdet_create(C.byref(cfg), self.gpu_id, C.byref(handle)))
    self._engine = handle
    self._finalized = False
    self._tensors = {}

    # image batch input [B, H, W, 3] fp32 (BGR, mean-subtracted)  (nn_skeleton.py:81-84)
    self.image_input = self.ph_image_input = GraphTensor(
        self, 'image_input', 0,
        (mc.BATCH_SIZE, mc.IMAGE_HEIGHT, mc.IMAGE_WIDTH, 3))
    self._tensors['image_input'] = self.image_input

    self.model_params = []           # GraphParam list, reference order
    self.model_size_counter = []     # (layer_name, parameter count)
    self.flop_counter = []           # (layer_name, flops)
    self.activation_counter = [('input', mc.IMAGE_WIDTH * mc.IMAGE_HEIGHT * 3)]
    self.preds = None
    self.det_boxes = self.det_probs = self.det_class = None

  # -------------------------------------------------------------------------------------
  def __del__(self):
    try:
      if getattr(self, '_engine', None):
        self._lib.sqdet_destroy(self._engine)
        self._engine = None
    except Exception:
      pass

  def _add_forward_graph(self):
    """NN architecture specification."""
    raise NotImplementedError

  def _new_tensor(self, name, tid):
    shape = (C.c_int64 * 4)()
    buf = C.create_string_buffer(256)
    _lib.check(self._lib.sqdet_tensor_info(self._engine, tid, buf, 256, shape))
    t = GraphTensor(self, buf.value.decode(), tid, tuple(int(s) for s in shape))
    self._tensors[t.name] = t
    return t

  def _count(self, layer_name, channels, filters, size, out_shape, relu):
    # parameter / flop / activation counters, nn_skeleton.py:549-561
    self.model_size_counter.append(
        (layer_name, (1 + size * size * int(channels)) * filters))
    num_flops = (1 + 2 * int(channels) * size * size) * filters * out_shape[1] * out_shape[2]
    if relu:
      num_flops += 2 * filters * out_shape[1] * out_shape[2]
    self.flop_counter.append((layer_name, num_flops))
    self.activation_counter.append(
        (layer_name, out_shape[1] * out_shape[2] * out_shape[3]))

  # ---- layer constructors -----------------------------------------------------------
  def _conv_layer(self, layer_name, inputs, filters, size, stride, padding='SAME',
                  freeze=False, xavier=False, relu=True, stddev=0.001):
    """Convolutional layer: relu?(conv2d(inputs) + biases)  (nn_skeleton.py:471-563).
    `freeze`, `xavier`, `stddev` only affect training/initialisation and are
    accepted for signature compatibility."""
    out = C.c_int()
    _lib.check(self._lib.sqdet_add_conv(
        self._engine, layer_name.encode(), inputs.id, int(filters), int(size),
        int(stride), _lib.pad_code(padding), int(bool(relu)), C.byref(out)))
    t = self._new_tensor(layer_name, out.value)
    channels = inputs.shape[3]
    self.model_params += [
        GraphParam(self, layer_name + '/kernels', (size, size, channels, filters)),
        GraphParam(self, layer_name + '/biases', (filters,))]
    self._count(layer_name, channels, filters, size, t.shape, relu)
    return t

  def _conv_bn_layer(self, inputs, conv_param_name, bn_param_name, scale_param_name,
                     filters, size, stride, padding='SAME', freeze=False, relu=True,
                     conv_with_bias=False, stddev=0.001):
    """Convolution + frozen BatchNorm + [relu]  (nn_skeleton.py:374-468).
    Variables live under scope `conv_param_name`: kernels, [biases], gamma, beta,
    mean, var (the Caffe bn_/scale_ names only matter when importing a .pkl)."""
    if padding.upper() != 'SAME':
      raise ValueError('_conv_bn_layer is only used with SAME padding')
    out = C.c_int()
    _lib.check(self._lib.sqdet_add_conv_bn(
        self._engine, conv_param_name.encode(), inputs.id, int(filters), int(size),
        int(stride), int(bool(relu)), int(bool(conv_with_bias)), C.byref(out)))
    t = self._new_tensor(conv_param_name, out.value)
    channels = inputs.shape[3]
    names = ['kernels'] + (['biases'] if conv_with_bias else []) + \
        ['gamma', 'beta', 'mean', 'var']
    for n in names:
      shp = (size, size, channels, filters) if n == 'kernels' else (filters,)
      self.model_params.append(GraphParam(self, conv_param_name + '/' + n, shp))
    self._count(conv_param_name, channels, filters, size, t.shape, relu)
    return t

  def _pooling_layer(self, layer_name, inputs, size, stride, padding='SAME'):
    """Max pooling (nn_skeleton.py:565-586)."""
    out = C.c_int()
    _lib.check(self._lib.sqdet_add_pool(
        self._engine, layer_name.encode(), inputs.id, int(size), int(stride),
        _lib.pad_code(padding), C.byref(out)))
    t = self._new_tensor(layer_name, out.value)
    self.activation_counter.append((layer_name, int(np.prod(t.shape[1:]))))
    return t

  def _fused_fire(self, layer_name, inputs, s1x1, e1x1, e3x3):
    """squeeze1x1 -> (expand1x1 || expand3x3) -> concat as ONE plan op so the engine
    can run the expand pair as a single fused tensor-core kernel
    (reference: three _conv_layer calls + tf.concat, squeezeDet.py:96-106)."""
    out = C.c_int()
    _lib.check(self._lib.sqdet_add_fire(
        self._engine, layer_name.encode(), inputs.id, int(s1x1), int(e1x1),
        int(e3x3), C.byref(out)))
    t = self._new_tensor(layer_name, out.value)
    cin = inputs.shape[3]
    sq_shape = t.shape[:3] + (s1x1,)
    for sub, ch, flt, sz in (('/squeeze1x1', cin, s1x1, 1),
                             ('/expand1x1', s1x1, e1x1, 1),
                             ('/expand3x3', s1x1, e3x3, 3)):
      nm = layer_name + sub
      self.model_params += [GraphParam(self, nm + '/kernels', (sz, sz, ch, flt)),
                            GraphParam(self, nm + '/biases', (flt,))]
      self._count(nm, ch, flt, sz, (sq_shape if sub == '/squeeze1x1'
                                    else t.shape[:3] + (flt,)), True)
    return t

  def _add_relu(self, name, a, b):
    """tf.nn.relu(a + b) of a residual unit (resnet50_convDet.py:55)."""
    out = C.c_int()
    _lib.check(self._lib.sqdet_add_add_relu(self._engine, name.encode(), a.id, b.id,
                                            C.byref(out)))
    return self._new_tensor(name, out.value)

  def _dropout(self, inputs, keep_prob, name=None):
    """tf.nn.dropout with keep_prob == 1.0 at inference: the identity."""
    assert keep_prob == 1.0
    return inputs

  # ---- interpretation graph ---------------------------------------------------------
  def _add_interpretation_graph(self):
    """Interpret NN output (nn_skeleton.py:142-283): declares `preds`/ANCHOR_BOX to
    the engine, freezes the plan and exposes det_boxes / det_probs / det_class."""
    mc = self.mc
    anchors = np.ascontiguousarray(np.asarray(mc.ANCHOR_BOX, dtype=np.float64))
    if anchors.ndim != 2 or anchors.shape[1] != 4:
      raise ValueError('mc.ANCHOR_BOX must be [ANCHORS, 4]')
    _lib.check(self._lib.sqdet_set_preds(
        self._engine, self.preds.id, anchors.ctypes.data, anchors.shape[0]))
    _lib.check(self._lib.sqdet_finalize(self._engine))
    self._finalized = True
    means = np.ascontiguousarray(np.asarray(mc.BGR_MEANS, dtype=np.float64).reshape(3))
    _lib.check(self._lib.sqdet_set_bgr_means(self._engine, means.ctypes.data))
    B, A = mc.BATCH_SIZE, anchors.shape[0]
    self.det_boxes = GraphTensor(self, 'bbox', None, (B, A, 4), 'det_boxes')
    self.det_probs = GraphTensor(self, 'score', None, (B, A), 'det_probs')
    self.det_class = GraphTensor(self, 'class_idx', None, (B, A), 'det_class')
    md = C.c_int32()
    _lib.check(self._lib.sqdet_results_dev(self._engine, None, None, None, None, None,
                                           C.byref(md)))
    self.max_dets = int(md.value)

  def _add_loss_graph(self):
    raise NotImplementedError('training graph: out of scope of the inference engine')

  _add_train_graph = _add_viz_graph = _add_loss_graph

  # ---- parameters -------------------------------------------------------------------
  def param_names(self):
    return [p.name for p in self.model_params]

  def set_param(self, name, value):
    arr = np.ascontiguousarray(np.asarray(value, dtype=np.float32))
    shape = (C.c_int64 * max(arr.ndim, 1))(*arr.shape)
    _lib.check(self._lib.sqdet_set_param(self._engine, name.encode(), arr.ctypes.data,
                                         shape, arr.ndim))

  def load_weights(self, weights, strict=True):
    """`weights`: mapping reference-variable-name -> ndarray (e.g. an .npz).  Stands
    in for tf.train.Saver(model.model_params).restore (demo.py:181-184)."""
    names = self.param_names()
    missing = [n for n in names if n not in weights]
    if strict and missing:
      raise KeyError('missing parameters: %s' % missing[:5])
    for n in names:
      if n in weights:
        self.set_param(n, weights[n])
    return missing

  # ---- execution ----------------------------------------------------------------------
  def _images_array(self, images):
    mc = self.mc
    arr = np.asarray(images, dtype=np.float32)
    want = (mc.BATCH_SIZE, mc.IMAGE_HEIGHT, mc.IMAGE_WIDTH, 3)
    if arr.shape != want:
      # TF raises ValueError on a feed-shape mismatch against the static placeholder
      raise ValueError('Cannot feed value of shape %r for image_input, which has '
                       'shape %r' % (arr.shape, want))
    return np.ascontiguousarray(arr)

  def detect(self, images, want_dets=False):
    """images [B,H,W,3] float -> (det_boxes [B,A,4] f32, det_probs [B,A] f32,
    det_class [B,A] i64) [+ (dets, counts) when want_dets]."""
    if not self._finalized:
      raise SqdetError(-4, 'model graph is not finalized')
    arr = self._images_array(images)
    B, A = self.det_probs.shape
    boxes = np.empty((B, A, 4), np.float32)
    probs = np.empty((B, A), np.float32)
    cls = np.empty((B, A), np.int64)
    dets = np.empty((B, self.max_dets), _lib.DET_DTYPE) if want_dets else None
    counts = np.empty((B,), np.int32) if want_dets else None
    _lib.check(self._lib.sqdet_detect(
        self._engine, arr.ctypes.data, boxes.ctypes.data, probs.ctypes.data,
        cls.ctypes.data, dets.ctypes.data if want_dets else None,
        counts.ctypes.data if want_dets else None, None))
    if want_dets:
      return boxes, probs, cls, dets, counts
    return boxes, probs, cls

  def detect_records(self, images):
    """images -> (dets [B,max_dets] structured, counts [B]) — only the filtered
    records cross PCIe (the end-to-end fast path)."""
    arr = self._images_array(images)
    B = self.mc.BATCH_SIZE
    dets = np.empty((B, self.max_dets), _lib.DET_DTYPE)
    counts = np.empty((B,), np.int32)
    _lib.check(self._lib.sqdet_detect(self._engine, arr.ctypes.data, None, None, None,
                                      dets.ctypes.data, counts.ctypes.data, None))
    return dets, counts

  # pipelined host path: copy of batch i+1 overlaps the compute of batch i (depth 2)
  def submit(self, images_ptr, dets_ptr, counts_ptr, img_type=_lib.IMG_F32):
    """Enqueue one batch (raw host pointers, ideally pinned; must stay valid until the
    matching wait()).  img_type: _lib.IMG_F32 (feed_dict semantics) or _lib.IMG_U8 (uint8
    BGR as cv2 returns it; `- mc.BGR_MEANS` happens on the GPU, demo.py:187-190)."""
    _lib.check(self._lib.sqdet_submit(self._engine, images_ptr, int(img_type), dets_ptr,
                                      counts_ptr))

  def wait(self):
    _lib.check(self._lib.sqdet_wait(self._engine))

  def detect_u8(self, images_u8):
    """uint8 BGR images [B,H,W,3] (already at mc.IMAGE_WIDTH x IMAGE_HEIGHT) ->
    (dets, counts): the demo.py:187-199 loop body from `im - mc.BGR_MEANS` on, on the GPU."""
    mc = self.mc
    arr = np.ascontiguousarray(np.asarray(images_u8, dtype=np.uint8))
    want = (mc.BATCH_SIZE, mc.IMAGE_HEIGHT, mc.IMAGE_WIDTH, 3)
    if arr.shape != want:
      raise ValueError('Cannot feed value of shape %r for image_input, which has shape %r'
                       % (arr.shape, want))
    dets = np.empty((mc.BATCH_SIZE, self.max_dets), _lib.DET_DTYPE)
    counts = np.empty((mc.BATCH_SIZE,), np.int32)
    self.submit(arr.ctypes.data, dets.ctypes.data, counts.ctypes.data, _lib.IMG_U8)
    self.wait()
    return dets, counts

  # ---- variable-size uint8 frames: pre-processing on the GPU (demo.py:187-190 / imdb.py:85-97) ----
  def submit_frames(self, frames, dets_ptr, counts_ptr, order='demo', rescale=False):
    """frames: list of B uint8 BGR arrays [h_i, w_i, 3] as cv2.imread returns them.  The engine
    resizes (cv2 float32 INTER_LINEAR) to (mc.IMAGE_WIDTH, mc.IMAGE_HEIGHT) and subtracts
    mc.BGR_MEANS in the reference order ('demo': resize then subtract, demo.py:187-190; 'eval':
    subtract then resize, imdb.py:85-97).  rescale=True: boxes are divided by each frame's
    (x_scale, y_scale) BEFORE filter_prediction, like eval.py:80-87.  Same wait() contract as
    submit(); the frame arrays must stay alive until then."""
    B = self.mc.BATCH_SIZE
    if len(frames) != B:
      raise ValueError('need %d frames, got %d' % (B, len(frames)))
    arrs = []
    for f in frames:
      a = np.ascontiguousarray(np.asarray(f, dtype=np.uint8))
      if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError('a frame must be uint8 [h, w, 3], got %r' % (a.shape,))
      arrs.append(a)
    ptrs = (C.c_void_p * B)(*[a.ctypes.data for a in arrs])
    hs = (C.c_int32 * B)(*[a.shape[0] for a in arrs])
    ws = (C.c_int32 * B)(*[a.shape[1] for a in arrs])
    code = {'demo': 0, 'eval': 1}[order]
    self._frames_alive = arrs
    _lib.check(self._lib.sqdet_submit_frames(self._engine, ptrs, hs, ws, code,
                                             int(bool(rescale)), dets_ptr, counts_ptr))

  def detect_frames(self, frames, order='demo', rescale=False):
    """Synchronous convenience over submit_frames: -> (dets [B,max_dets], counts [B])."""
    B = self.mc.BATCH_SIZE
    dets = np.empty((B, self.max_dets), _lib.DET_DTYPE)
    counts = np.empty((B,), np.int32)
    self.submit_frames(frames, dets.ctypes.data, counts.ctypes.data, order, rescale)
    self.wait()
    return dets, counts

  def set_box_scale(self, scales):
    """eval.py:83-84 for host-resized inputs: `scales` = B (x_scale, y_scale) pairs, or None to
    switch the rescale off.  Later forwards divide det_boxes by them before the filter."""
    if scales is None:
      _lib.check(self._lib.sqdet_set_box_scale(self._engine, None))
      return
    arr = np.ascontiguousarray(np.asarray(scales, dtype=np.float32).reshape(-1))
    if arr.size != 2 * self.mc.BATCH_SIZE:
      raise ValueError('need BATCH_SIZE (x_scale, y_scale) pairs')
    _lib.check(self._lib.sqdet_set_box_scale(self._engine, arr.ctypes.data))

  # ---- multi-GPU: the one all-gather of the filtered records (shard.py drives this) ------------
  def comm_init(self, nranks, rank, unique_id, in_forward=True):
    """ncclCommInitRank on this engine's device; `unique_id` = the 128 bytes rank 0 obtained
    from _lib.comm_unique_id() and shared through any host channel.  in_forward=True makes
    every forward end with the all-gather (captured in its CUDA graph)."""
    buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
    _lib.check(self._lib.sqdet_comm_init(self._engine, int(nranks), int(rank), buf))
    if in_forward:
      _lib.check(self._lib.sqdet_set_gather_in_forward(self._engine, 1))

  def set_gather_in_forward(self, on):
    _lib.check(self._lib.sqdet_set_gather_in_forward(self._engine, int(bool(on))))

  def allgather(self, stream=None):
    _lib.check(self._lib.sqdet_allgather(self._engine, None, stream))

  def gathered_device(self):
    """(device pointer, bytes per rank, nranks) of the all-gather receive buffer."""
    p, nb, nr = C.c_void_p(), C.c_int64(), C.c_int32()
    _lib.check(self._lib.sqdet_gathered_dev(self._engine, C.byref(p), C.byref(nb), C.byref(nr)))
    return p.value, int(nb.value), int(nr.value)

  def read_gathered(self):
    """Host copy of the receive buffer as [nranks, bytes_per_rank] uint8 (synchronous)."""
    ptr, nb, nr = self.gathered_device()
    out = np.empty((nr, nb), np.uint8)
    _lib.check(self._lib.sqdet_stream_sync(self.gpu_id, self.engine_stream()))
    _lib.check(self._lib.sqdet_memcpy_d2h(out.ctypes.data, ptr, out.nbytes, None))
    _lib.check(self._lib.sqdet_stream_sync(self.gpu_id, None))
    return out

  def comm_destroy(self):
    _lib.check(self._lib.sqdet_comm_destroy(self._engine))

  def engine_stream(self):
    return self._lib.sqdet_engine_stream(self._engine)

  @staticmethod
  def records_to_lists(dets, count):
    """One image's records -> the reference's filter_prediction return triple."""
    if count < 0:
      raise SqdetError(-6, 'more boxes above PROB_THRESH than the record capacity')
    d = dets[:count]
    boxes = [np.array([r['cx'], r['cy'], r['w'], r['h']], dtype=np.float32) for r in d]
    probs = [np.float32(r['prob']) for r in d]
    cls = [int(r['cls']) for r in d]
    return boxes, probs, cls

  def detect_filtered(self, images):
    """detect + filter_prediction for every image: list of (final_boxes,
    final_probs, final_cls_idx) per image (demo.py:193-199 in one GPU pass)."""
    dets, counts = self.detect_records(images)
    return [self.records_to_lists(dets[i], int(counts[i])) for i in range(len(counts))]

  def run(self, fetches, feed_dict):
    """sess.run equivalent for fetches among det_boxes/det_probs/det_class/preds/any
    activation handle."""
    if self.image_input not in feed_dict and self.ph_image_input not in feed_dict:
      raise ValueError('feed_dict must feed model.image_input')
    images = feed_dict[self.image_input]
    boxes, probs, cls = self.detect(images)
    table = {'det_boxes': boxes, 'det_probs': probs, 'det_class': cls}
    out = []
    for f in fetches:
      if f.fetch is not None:
        out.append(table[f.fetch])
      else:
        out.append(self.read_tensor(f))
    return out

  def read_tensor(self, t):
    """Fetch an activation of the last forward (debug / layer-wise parity)."""
    if isinstance(t, str):
      t = self._tensors[t]
    out = np.empty(t.shape, np.float32)
    _lib.check(self._lib.sqdet_read_tensor(self._engine, t.id, out.ctypes.data))
    return out

  # device-resident path (no host copies; used by bench / multi-GPU runner)
  def forward_device(self, images_dev_ptr, stream=None):
    _lib.check(self._lib.sqdet_forward(self._engine, images_dev_ptr, stream))

  def forward_profiled(self, images_dev_ptr, stream=None):
    n = self._lib.sqdet_num_ops(self._engine)
    ms = np.zeros(n, np.float32)
    _lib.check(self._lib.sqdet_forward_profiled(self._engine, images_dev_ptr, stream,
                                                ms.ctypes.data))
    return list(zip(self.op_table(), ms.tolist()))

  def op_table(self):
    """[(name, flops, params, min_bytes)] per plan op (+ the two post-proc ops)."""
    rows = []
    for i in range(self._lib.sqdet_num_ops(self._engine)):
      buf = C.create_string_buffer(256)
      fl, pa, by = C.c_int64(), C.c_int64(), C.c_int64()
      _lib.check(self._lib.sqdet_op_info(self._engine, i, buf, 256, C.byref(fl),
                                         C.byref(pa), C.byref(by)))
      rows.append((buf.value.decode(), fl.value, pa.value, by.value))
    return rows

  def results_device(self):
    """Device pointers of the last forward's results: dict of ints + max_dets."""
    ptrs = [C.c_void_p() for _ in range(5)]
    md = C.c_int32()
    _lib.check(self._lib.sqdet_results_dev(self._engine, *[C.byref(p) for p in ptrs],
                                           C.byref(md)))
    keys = ('det_boxes', 'det_probs', 'det_class', 'dets', 'counts')
    out = {k: p.value for k, p in zip(keys, ptrs)}
    out['max_dets'] = int(md.value)
    return out

  def launches_per_forward(self):
    return self._lib.sqdet_launches_per_forward(self._engine)

  # ---- filter_prediction ---------------------------------------------------------------
  def filter_prediction(self, boxes, probs, cls_idx):
    """Filter bounding box predictions with probability threshold and non-maximum
    suppression (nn_skeleton.py:696-734) — same arguments and return triple, computed
    by the GPU filter kernel through `sqdet_topk_nms`.

    Args:
      boxes: array of [cx, cy, w, h].   probs: array of probabilities.
      cls_idx: array of class indices.
    Returns:
      final_boxes (list of ndarray(4)), final_probs (list of float32),
      final_cls_idx (list of int).
    """
    mc = self.mc
    boxes = np.ascontiguousarray(np.asarray(boxes, dtype=np.float32)).reshape(-1, 4)
    probs = np.ascontiguousarray(np.asarray(probs, dtype=np.float32)).reshape(-1)
    cls_idx = np.ascontiguousarray(np.asarray(cls_idx, dtype=np.int64)).reshape(-1)
    n = len(probs)
    if n == 0:
      return [], [], []
    topn = 0 < mc.TOP_N_DETECTION < n
    cap = int(mc.TOP_N_DETECTION) if topn else min(n, 1024)
    # one pooled, grow-only scratch allocation instead of five cudaMalloc/cudaFree pairs per call
    pool = _lib.scratch_pool(self.gpu_id)
    p_boxes, p_probs, p_cls, p_dets, p_cnt = pool.carve(
        boxes.nbytes, probs.nbytes, cls_idx.nbytes, cap * _lib.DET_DTYPE.itemsize, 4)
    pool.upload(p_boxes, boxes)
    pool.upload(p_probs, probs)
    pool.upload(p_cls, cls_idx)
    _lib.check(self._lib.sqdet_topk_nms(
        p_boxes, p_probs, p_cls, 1, n, int(mc.CLASSES), int(mc.TOP_N_DETECTION),
        float(mc.PROB_THRESH), float(mc.NMS_THRESH), p_dets, p_cnt, cap, None))
    dets = pool.download(p_dets, _lib.DET_DTYPE, (cap,))
    count = int(pool.download(p_cnt, np.int32, (1,))[0])
    return self.records_to_lists(dets, count)
