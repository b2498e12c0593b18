# Reference-side adapter (NOT run in this repository: no R in the image; see INTEGRATION.md).
# Added to const-ae/mixdir as R/mixdir_vi_b200.R.  Same arguments and same result list as
# mixdir_vi() (R/mixdir_vi.R:2-3,143-155) and mixdir_vi_dp() (R/mixdir_vi_dp.R:2-3,172-185).

mixdir_vi_b200 <- function(X, n_latent, alpha, beta, categories, max_iter, epsilon,
                           select_latent=FALSE, omega_init=NULL, kappa1_init=NULL, kappa2_init=NULL,
                           zeta_init=NULL, phi_init=NULL, verbose=FALSE, devices=integer(0)){
  n_ind <- nrow(X); n_quest <- ncol(X)
  n_cat <- max(X, na.rm=TRUE)
  # defaults exactly as R/mixdir_vi.R:10-26 / R/mixdir_vi_dp.R:10-29
  if(is.null(alpha) || length(alpha) == 0){ alpha1 <- 1; alpha2 <- 1
  }else if(select_latent && length(alpha) == 2){ alpha1 <- alpha[1]; alpha2 <- alpha[2]
  }else{
    if(length(alpha) > 1 || select_latent)
      warning(paste0("alpha should only be a single value. Using alpha=alpha[1]=", alpha[1]))
    alpha1 <- alpha[1]; alpha2 <- alpha[1]
  }
  if(is.null(beta) || length(beta) == 0) beta <- 0.1 else beta <- beta[1]
  # random initial values stay in R (R/mixdir_vi.R:31-39)
  if(is.null(zeta_init)) zeta_init <- extraDistr::rdirichlet(n_ind, rep(1, n_latent))
  if(is.null(phi_init)){
    phi_init <- lapply(1:n_quest, function(j) lapply(1:n_latent, function(k)
      sample(1:3, size=length(categories[[j]]), replace=TRUE)))
  }else{
    for(k in 1:n_latent){    # the shape check of R/mixdir_vi.R:41-47
      len <- sapply(phi_init, function(x) length(x[[k]]))
      if(! all(len == sapply(categories, length)))
        stop(paste0("phi_init has the wrong number of elements for feature: ",
                    paste0(which(len != sapply(categories, length)), collapse = ", ")))
    }
  }
  res <- mixdir_vi_b200_cpp(X, as.integer(sapply(categories, length)), n_latent, select_latent,
                            alpha1, alpha2, beta, max_iter, epsilon, colSums(zeta_init),
                            as.numeric(unlist(phi_init)), n_cat, as.integer(devices))
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

# The selector in mixdir() -- a three-line change at R/mixdir.R:121-128:
#
#   engine <- getOption("mixdir.engine", "cpu")            # default: the reference CPU path
#   result_tmp <- if(engine == "b200")
#       mixdir_vi_b200(X=X, n_latent=n_latent, alpha=alpha, beta=beta, categories=categories,
#                      max_iter=max_iter, epsilon=epsilon, select_latent=select_latent, ...)
#     else if(! select_latent) mixdir_vi(...) else mixdir_vi_dp(...)
