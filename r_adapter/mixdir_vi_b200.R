# Reference-side adapter (NOT run in this repository: no R in the image; see INTEGRATION.md).
# Added to const-ae/mixdir as R/mixdir_vi_b200.R.  Same arguments and same result list as
# mixdir_vi() (R/mixdir_vi.R:2-3,143-155) and mixdir_vi_dp() (R/mixdir_vi_dp.R:2-3,172-185), plus
# `repetitions`: the restarts of one mixdir() call (R/mixdir.R:120-134) share ONE upload of X and
# run as one library call; the result with the largest ELBO is returned.

mixdir_vi_b200 <- function(X, n_latent, alpha, beta, categories, max_iter, epsilon,
                           select_latent=FALSE, repetitions=1, zeta_init=NULL, phi_init=NULL,
                           verbose=FALSE, devices=integer(0), handle=NULL){
  n_ind <- nrow(X); n_quest <- ncol(X)
  n_cat <- max(X, na.rm=TRUE)
  ncat <- as.integer(sapply(categories, length))
  # defaults exactly as R/mixdir_vi.R:10-26 / R/mixdir_vi_dp.R:10-29
  if(is.null(alpha) || length(alpha) == 0){ alpha1 <- 1; alpha2 <- 1
  }else if(select_latent && length(alpha) == 2){ alpha1 <- alpha[1]; alpha2 <- alpha[2]
  }else{
    if(length(alpha) > 1 || select_latent)
      warning(paste0("alpha should only be a single value. Using alpha=alpha[1]=", alpha[1]))
    alpha1 <- alpha[1]; alpha2 <- alpha[1]
  }
  if(is.null(beta) || length(beta) == 0) beta <- 0.1 else beta <- beta[1]
  # random initial values stay in R, one draw per repetition (R/mixdir_vi.R:31-39)
  colsum_init <- matrix(0, nrow=n_latent, ncol=repetitions)
  phi_flat <- matrix(0, nrow=n_latent * sum(ncat), ncol=repetitions)
  for(repit in seq_len(repetitions)){
    z0 <- if(is.null(zeta_init)) extraDistr::rdirichlet(n_ind, rep(1, n_latent)) else zeta_init
    p0 <- phi_init
    if(is.null(p0)){
      p0 <- lapply(1:n_quest, function(j) lapply(1:n_latent, function(k)
        sample(1:3, size=length(categories[[j]]), replace=TRUE)))
    }else{
      for(k in 1:n_latent){    # the shape check of R/mixdir_vi.R:41-47
        len <- sapply(p0, function(x) length(x[[k]]))
        if(! all(len == ncat))
          stop(paste0("phi_init has the wrong number of elements for feature: ",
                      paste0(which(len != ncat), collapse = ", ")))
      }
    }
    colsum_init[, repit] <- colSums(z0)
    phi_flat[, repit] <- as.numeric(unlist(p0))
  }
  own <- is.null(handle)
  if(own) handle <- upload_b200_cpp(X, ncat, as.integer(devices))   # one upload for all repetitions
  res <- tryCatch(
    mixdir_vi_b200_cpp(handle, ncat, n_latent, select_latent, alpha1, alpha2, beta, max_iter,
                       epsilon, colsum_init, phi_flat, n_cat),
    finally = if(own) release_b200_cpp(handle))
  # ragged (j,k,r) vectors -> the reference's nested named lists (R/mixdir_vi.R:119-130)
  unflatten <- function(flat){
    o <- 0
    out <- lapply(1:n_quest, function(j){
      cj <- length(categories[[j]])
      l <- lapply(1:n_latent, function(k){
        x <- flat[o + (k-1)*cj + 1:cj]; names(x) <- categories[[j]]; x })
      o <<- o + n_latent * cj
      l })
    out
  }
  U <- unflatten(res$U_flat); names(U) <- colnames(X)
  phi <- lapply(unflatten(res$phi_flat), function(l) lapply(l, unname))
  if(verbose) for(i in seq_along(res$convergence)) if(i %% 10 == 0)
    message(paste0("Iter: ", i, " ELBO: ", formatC(res$convergence[i], digits=4)))
  specific <- if(select_latent)
    list(kappa1=res$prior[1:n_latent], kappa2=res$prior[n_latent + 1:n_latent], phi=phi)
  else list(omega=res$prior, phi=phi)
  list(converged=res$converged, convergence=res$convergence, ELBO=res$ELBO, lambda=res$lambda,
       pred_class=res$pred_class, class_prob=res$class_prob, category_prob=U,
       specific_params=specific)
}

# The selector in mixdir() replaces the repetitions loop at R/mixdir.R:120-134:
#
#   if(getOption("mixdir.engine", "cpu") == "b200"){
#     result <- mixdir_vi_b200(X=X, n_latent=n_latent, alpha=alpha, beta=beta, categories=categories,
#                              max_iter=max_iter, epsilon=epsilon, select_latent=select_latent,
#                              repetitions=repetitions, ...)
#   }else{ <the existing loop> }


# predict_class() (R/predict.R:39-105) with the arithmetic of :88-103 on the device.  Everything
# up to the coded matrix is the reference's own code (name matching, "(Missing)", the warning for
# unseen levels); only the coding to integers is new, because the device wants codes, not strings.
predict_class_b200 <- function(X, lambda, category_prob, na_handle=c("ignore", "category")){
  categories <- lapply(category_prob, function(cat) names(cat[[1]]))
  if(is.null(na_handle)) na_handle <- "ignore" else na_handle <- match.arg(na_handle, c("ignore", "category"))
  if(is.vector(X) || is.list(X)){
    tmp <- as.data.frame(as.list(X), stringsAsFactors = FALSE); colnames(tmp) <- names(X); X <- tmp
  }else if(is.matrix(X)){
    tmp <- as.data.frame(X, stringsAsFactors = FALSE); colnames(tmp) <- colnames(X); X <- tmp
  }
  n_ind <- nrow(X)
  codes <- vapply(names(categories), function(coln){
    vec <- X[[coln]]
    if(is.null(vec)) return(rep(NA_real_, n_ind))            # R/predict.R:66-67
    vec <- as.character(vec)
    if(na_handle == "category") vec[is.na(vec)] <- "(Missing)"
    pos_cats <- categories[[coln]]
    non_matches <- which(!vec %in% pos_cats & ! is.na(vec))
    if(length(non_matches) > 0){                             # R/predict.R:74-80
      warning(paste0("The new data has a response (",
                     paste0("\"", unique(vec[non_matches]), "\"", collapse = ","),
                     ") for category \"", coln, "\" which is not in category_prob. ",
                     "Handling X[j, i] it as if it was missing.\n"))
      vec[non_matches] <- NA
    }
    as.numeric(match(vec, pos_cats))                         # the index category_prob[[j]][[k]][X[,j]] uses
  }, FUN.VALUE=rep(0.0, times=n_ind))
  codes <- matrix(codes, nrow=n_ind)
  handle <- upload_b200_cpp(codes, as.integer(sapply(categories, length)), integer(0))
  tryCatch(predict_class_b200_cpp(handle, lambda, as.numeric(unlist(category_prob))),
           finally = release_b200_cpp(handle))
}


# find_defining_features() (R/predict.R:266-339): the greedy loop stays as it is; the `vapply` over
# the remaining variables (:299-310), one predict.mixdir() call each, becomes one device pass per
# round on a handle that holds the coded subsample for the whole search.
#
#   handle <- upload_b200_cpp(X_coded[subsample, ], ncat, integer(0))
#   while(...){
#     loss <- feature_divergence_b200_cpp(handle, mixdir_obj$lambda,
#                                         as.numeric(unlist(mixdir_obj$category_prob)),
#                                         colnames(X) %in% remaining, measure == "ARI")
#     quality_loss <- loss[match(remaining, colnames(X))]     # replaces the vapply
#     ...
#   }
#   final_quality <- loss[ncol(X) + 1]                        # R/predict.R:326-331
#   release_b200_cpp(handle)
