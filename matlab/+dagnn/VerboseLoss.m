classdef VerboseLoss < dagnn.Loss
%VERBOSELOSS  dagnn.Loss under the name the reference uses.
%   SOURCE-ONLY: there is no MATLAB in the build image, so this file has never been executed.
%   emoVoxCeleb/emoVoxZoo.m:160-163 attaches dagnn.VerboseLoss('loss', 'classerror') on {'prediction', 'maxLabel'} ->
%   'classerror'; emoVoxCeleb/run_distillation.m:200-203 reads its .ignoreAverage and .average exactly as for a
%   dagnn.Loss.  The class is not part of public MatConvNet / mcnExtraLayers; everything the scripts rely on is
%   inherited from dagnn.Loss (forward = vl_nnloss(x, c, [], 'loss', obj.loss), running .average, .numAveraged).
%   The fused device path accumulates the same class error in the loss kernel (scalars[1]).

  properties
    verbose = false   % print the running average after every batch
  end

  methods
    function obj = VerboseLoss(varargin)
      obj.load(varargin) ;
    end

    function outputs = forward(obj, inputs, params)
      outputs = forward@dagnn.Loss(obj, inputs, params) ;
      if obj.verbose
        fprintf('%s: %.4f (running average over %d)\n', obj.loss, obj.average, obj.numAveraged) ;
      end
    end
  end
end
