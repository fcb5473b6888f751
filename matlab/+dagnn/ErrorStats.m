classdef ErrorStats < dagnn.Loss
%ERRORSTATS  Per-class running accuracy metric layer.
%   SOURCE-ONLY: there is no MATLAB in the build image, so this file has never been executed.  It restates, for a
%   MATLAB host, the metric block that emoVoxCeleb/emoVoxZoo.m:166-169 attaches as
%       dagnn.ErrorStats('numClasses', numOutputs) on {'prediction', 'maxLabel'} -> 'classAccs'
%   and whose fields emoVoxCeleb/run_distillation.m:186-207 (extractStats) reads:
%       .average    1 x numClasses running accuracy of each class (samples of class k predicted as k / samples of class k)
%       .classDist  1 x numClasses number of samples seen per class
%   The class is not part of public MatConvNet / mcnExtraLayers.  The fused device path computes the same counters inside
%   the loss kernel (csrc/hbm_kernels.cuh: softmaxce_fused_kernel, class_stats = [correct | count]); the CPU restatement is
%   oracle/mcn_ops.py: error_stats.  No gradient flows through this block.

  properties
    numClasses = 8
  end

  properties (Transient)
    classDist = []     % samples seen per class since the last reset
    classCorrect = []  % correctly classified samples per class
  end

  methods
    function obj = ErrorStats(varargin)
      obj.load(varargin) ;
      obj.loss = 'classerror' ;
      obj.reset() ;
    end

    function outputs = forward(obj, inputs, params) %#ok<INUSD>
      x = gather(inputs{1}) ;                       % 1 x 1 x numClasses x N predictions
      c = gather(inputs{2}) ;                       % 1 x 1 x 1 x N labels, 1-based
      [~, pred] = max(x, [], 3) ;                   % first maximum wins (MatConvNet vl_nnloss 'classerror')
      pred = pred(:)' ; c = c(:)' ;
      for k = 1:obj.numClasses
        sel = (c == k) ;
        obj.classDist(k) = obj.classDist(k) + sum(sel) ;
        obj.classCorrect(k) = obj.classCorrect(k) + sum(pred(sel) == k) ;
      end
      seen = max(obj.classDist, 1) ;
      obj.average = obj.classCorrect ./ seen ;      % classes not seen yet report 0
      obj.numAveraged = sum(obj.classDist) ;
      outputs{1} = mean(obj.average) ;
    end

    function [derInputs, derParams] = backward(obj, inputs, params, derOutputs) %#ok<INUSD>
      derInputs = {[], []} ;                        % a metric: no gradient
      derParams = {} ;
    end

    function reset(obj)
      obj.average = zeros(1, obj.numClasses) ;
      obj.numAveraged = 0 ;
      obj.classDist = zeros(1, obj.numClasses) ;
      obj.classCorrect = zeros(1, obj.numClasses) ;
    end

    function outputSizes = getOutputSizes(obj, inputSizes) %#ok<INUSD>
      outputSizes{1} = [1 1 1 1] ;
    end
  end
end
